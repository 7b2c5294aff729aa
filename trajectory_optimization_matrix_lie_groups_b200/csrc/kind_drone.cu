// Kernel sequences of one problem family (explicit instantiation; see host_impl.cuh)
#include "host_impl.cuh"

TRAJOPT_KIND_INSTANTIATE(TRAJOPT_DRONE)
