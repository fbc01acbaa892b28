// TEST INFRASTRUCTURE ONLY (oracle/): never imported by the product path.
//
// Builds the reference's OWN CPU implementation of nms / roi_align_forward into
// oracle/_ref/ from the sources where they lie under /root/reference
// (lib/model/csrc/vision.cpp, cpu/nms_cpu.cpp, cpu/ROIAlign_cpu.cpp) -- no copy, no edit.
//
// The unmodified sources fail on torch >= 2.x at exactly two places
// (cpu/ROIAlign_cpu.cpp:242 and cpu/nms_cpu.cpp:71 pass `x.type()` -- a
// DeprecatedTypeProperties -- to AT_DISPATCH_FLOATING_TYPES).  Instead of patching a
// copy we re-define that one macro so it accepts DeprecatedTypeProperties, then
// #include the reference translation units verbatim (found through -I<ref>/lib/model/csrc).
#include <torch/extension.h>

static inline at::ScalarType aitref_scalar_type(const at::DeprecatedTypeProperties& t) {
  return t.scalarType();
}
static inline at::ScalarType aitref_scalar_type(at::ScalarType t) { return t; }

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(aitref_scalar_type(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))

#include "cpu/nms_cpu.cpp"
#include "cpu/ROIAlign_cpu.cpp"
#include "vision.cpp"   // PYBIND11_MODULE(TORCH_EXTENSION_NAME, m): nms, roi_align_forward, ...
