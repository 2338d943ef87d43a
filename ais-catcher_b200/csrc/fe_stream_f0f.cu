// fe_stream_f0f.cu -- streaming front end: CF32, two-warp CTAs (experiments).
#include "fe_stream.cuh"

namespace aisgpu {

template cudaError_t launch_frontend_stream_shape<0, 32, 3, 2>(const FeParams &, int, bool, long long, cudaStream_t);
template cudaError_t launch_frontend_stream_shape<0, 32, 4, 2>(const FeParams &, int, bool, long long, cudaStream_t);

} // namespace aisgpu
