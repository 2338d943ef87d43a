// fe_stream_f0b.cu -- streaming front end: CF32, ring of 4 chunks, one-warp CTAs; one translation unit per shape keeps the build parallel.
#include "fe_stream.cuh"

namespace aisgpu {

template cudaError_t launch_frontend_stream_shape<0, 4, 1>(const FeParams &, int, bool, long long, cudaStream_t);

} // namespace aisgpu
