#define SMF_K 3
#include "fb2m_launch.inc"
