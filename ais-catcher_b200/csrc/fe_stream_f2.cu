// fe_stream_f2.cu -- streaming front end: CS8; one translation unit per shape keeps the build parallel.
#include "fe_stream.cuh"

namespace aisgpu {

template cudaError_t launch_frontend_stream_shape<2, 16, 8, 1>(const FeParams &, int, bool, int, cudaStream_t);

} // namespace aisgpu
