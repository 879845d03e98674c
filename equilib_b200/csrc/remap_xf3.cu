#define EQB_XF 3
#include "remap_launch.inl"
