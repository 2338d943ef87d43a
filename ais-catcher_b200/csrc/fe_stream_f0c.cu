// fe_stream_f0c.cu -- streaming front end: CF32, ring of 8 chunks, four-warp CTAs (round-1 shape, kept for A/B runs); one translation unit per shape keeps the build parallel.
#include "fe_stream.cuh"

namespace aisgpu {

template cudaError_t launch_frontend_stream_shape<0, 8, 4>(const FeParams &, int, bool, long long, cudaStream_t);

} // namespace aisgpu
