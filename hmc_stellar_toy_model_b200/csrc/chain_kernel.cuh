// Warp-per-chain one-star kernel (placeholder until the specialised kernel lands; the CTA-per-field kernel
// handles every configuration).
#pragma once
#include "common.cuh"
namespace srhmc {
template <typename T> inline int configure_chain_kernel(const FieldParams&) { return (int)cudaErrorNotSupported; }
template <typename T> inline int launch_chain_kernel(const FieldParams&, const LaunchArgs&, double*, int, cudaStream_t) { return (int)cudaErrorNotSupported; }
}  // namespace srhmc
