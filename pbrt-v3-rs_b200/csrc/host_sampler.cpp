// Host-side sampler tables (Halton permutations) — filled in with the path tracer.
#include "../../include/b200pt.h"
