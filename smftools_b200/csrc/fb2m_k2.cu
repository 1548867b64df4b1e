#define SMF_K 2
#include "fb2m_launch.inc"
