#define SMF_K 3
#include "fb2w_launch.inc"
