// K = 2 only needs the boundary walk + any-length posterior sweeps (reads beyond the on-chip reach)
#define SMF_K 2
#include "fb2w_launch.inc"
