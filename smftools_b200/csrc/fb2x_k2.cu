// extra chunk lengths (21, 25, 29) of the K = 2 posterior kernel
#define SMF_K 2
#define SMF_FB2_EXTRA 1
#include "fb2_launch.inc"
