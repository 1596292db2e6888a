// Instantiations of the fused q-KG kernels for padded dimension 2 (one translation unit per dimension so the build
// parallelises); registered into the run-time dispatch table of kg.cu.
#include "kg_mc.cuh"

namespace cmoe {
namespace {
const KgDispatchEntry kEntries[] = {CMOE_KG_ENTRIES_FOR_DIM(2)};
struct Registrar {
  Registrar() { register_kg_entries(kEntries, static_cast<int>(sizeof(kEntries) / sizeof(kEntries[0]))); }
} registrar;
}  // namespace
}  // namespace cmoe
