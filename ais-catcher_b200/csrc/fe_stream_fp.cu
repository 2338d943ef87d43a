// fe_stream_fp.cu -- streaming front end for -go FP_DS on: CU8 through the reference's packed-uint16 integer CIC stages
// (DSP::Downsample16_CU8, DSP.cpp:499-665), then the float 96 kHz tail; only K = 4 (the exact 1536K bucket) exists.
#include "fe_stream.cuh"

namespace aisgpu {

cudaError_t launch_frontend_stream_fpds(const FeParams &p, int forced_L, cudaStream_t s) { return launch_st_one<4, 4, 16, 8, 1, false>(p, forced_L, s); }

} // namespace aisgpu
