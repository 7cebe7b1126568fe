// libsia_b200_debug.so: the bring-up probes of include/sia_b200_debug.h group (2) -- tcgen05 / TMA / TMEM / ALU
// micro-benchmarks.  Its own translation unit (own copy of the internal helpers, none of the product's exports),
// so the product library libsia_b200.so carries no probe code.
#define SIA_DEBUG_LIB 1
#include "core.cu"
#include "tmap.cu"
#include "probe.cu"
