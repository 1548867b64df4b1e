#define SMF_K 2
#include "fb2b_launch.inc"
