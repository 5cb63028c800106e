// oracle/ref_compat.h -- TEST INFRASTRUCTURE ONLY.
// Force-included (-include) when oracle/build_ref.py compiles the UNMODIFIED reference sources:
// restores the `detail::scalar_type(DeprecatedTypeProperties)` overload that
// AT_DISPATCH_FLOATING_TYPES(value.type(), ...) (reference cuda/ms_deform_attn_cuda.cu:64,134)
// relied on and that newer torch releases removed.
#pragma once
#ifdef __cplusplus
#include <ATen/ATen.h>
#include <ATen/core/DeprecatedTypeProperties.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
#endif
