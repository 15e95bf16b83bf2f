// explicit instantiation of the per-curve engine entry points
#include "curve_impl.cuh"
namespace b200 {
B200_INSTANTIATE(G1_377)
}
