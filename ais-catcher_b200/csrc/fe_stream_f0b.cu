// fe_stream_f0b.cu -- streaming front end: CF32, 64-sample chunks (512 bytes per lane and visit), ring of 3, four-warp CTAs; one translation unit per shape keeps the build parallel.
#include "fe_stream.cuh"

namespace aisgpu {

template cudaError_t launch_frontend_stream_shape<0, 64, 3, 4>(const FeParams &, int, bool, long long, cudaStream_t);

} // namespace aisgpu
