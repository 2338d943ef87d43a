// fe_stream_f0d.cu -- streaming front end: CF32, 32-sample chunks, ring of 5, four-warp CTAs; one translation unit per shape keeps the build parallel.
#include "fe_stream.cuh"

namespace aisgpu {

template cudaError_t launch_frontend_stream_shape<0, 32, 5, 4>(const FeParams &, int, bool, int, cudaStream_t);

} // namespace aisgpu
