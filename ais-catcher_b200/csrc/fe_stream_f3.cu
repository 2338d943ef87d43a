// fe_stream_f3.cu -- streaming front end: CS16; one translation unit per shape keeps the build parallel.
#include "fe_stream.cuh"

namespace aisgpu {

template cudaError_t launch_frontend_stream_shape<3, 16, 8, 1>(const FeParams &, int, bool, int, cudaStream_t);

} // namespace aisgpu
