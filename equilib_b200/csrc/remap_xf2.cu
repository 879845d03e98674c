#define EQB_XF 2
#include "remap_launch.inl"
