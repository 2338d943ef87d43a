// fe_stream_f0e.cu -- streaming front end: CF32 shapes with a shallower ring, so that two to four CTAs share an SM (experiments).
#include "fe_stream.cuh"

namespace aisgpu {

template cudaError_t launch_frontend_stream_shape<0, 32, 2, 4>(const FeParams &, int, bool, long long, cudaStream_t);
template cudaError_t launch_frontend_stream_shape<0, 32, 3, 4>(const FeParams &, int, bool, long long, cudaStream_t);

} // namespace aisgpu
