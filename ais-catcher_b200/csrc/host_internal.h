// Hooks the host-side units of libaisgpu.so (feeder.cpp) use besides the public ABI; defined in aisgpu.cu, not part of include/aisgpu.h.
#pragma once
#include "../../include/aisgpu.h"
extern "C" {
const aisgpu_config *aisgpu_internal_config(aisgpu_handle *h);
void aisgpu_internal_set_error(aisgpu_handle *h, const char *msg);
}
