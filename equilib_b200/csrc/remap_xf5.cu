#define EQB_XF 5
#include "remap_launch.inl"
