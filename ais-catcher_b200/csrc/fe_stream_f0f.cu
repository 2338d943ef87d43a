// fe_stream_f0f.cu -- streaming front end: CF32, 32-sample chunks, ring of 4, four-warp CTAs (139 KB).
#include "fe_stream.cuh"

namespace aisgpu {

template cudaError_t launch_frontend_stream_shape<0, 32, 4, 4>(const FeParams &, int, bool, int, cudaStream_t);

} // namespace aisgpu
