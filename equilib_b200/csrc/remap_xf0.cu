#define EQB_XF 0
#include "remap_launch.inl"
