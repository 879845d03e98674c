#define EQB_XF 1
#include "remap_launch.inl"
