#define EQB_XF 4
#include "remap_launch.inl"
