// fe_stream_f0c.cu -- streaming front end: CF32, 16-sample chunks, ring of 8, four-warp CTAs (the default shape); one translation unit per shape keeps the build parallel.
#include "fe_stream.cuh"

namespace aisgpu {

template cudaError_t launch_frontend_stream_shape<0, 16, 8, 4>(const FeParams &, int, bool, long long, cudaStream_t);

} // namespace aisgpu
