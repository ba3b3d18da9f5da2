// placeholder until the tcgen05 path lands
#include "gb_tc.h"
namespace gb {
TcGridWorkspace::~TcGridWorkspace() {
  for (auto p : x0) if (p) cudaFree(p);
  if (list_xyzr) cudaFree(list_xyzr);
  if (list_ch) cudaFree(list_ch);
  if (list_n) cudaFree(list_n);
}
void TcWorkspace::ensure(int i, size_t bytes) {
  if (cap[i] >= bytes) return;
  if (buf[i]) cudaFree(buf[i]);
  GB_CUDA(cudaMalloc(&buf[i], bytes));
  cap[i] = bytes;
}
TcWorkspace::~TcWorkspace() { for (auto p : buf) if (p) cudaFree(p); }
bool tc_supported(const Model&) { return false; }
int tc_forward(const Model&, const TcPoseBatch&, TcGridWorkspace&, TcWorkspace&, float*, cudaStream_t, Profiler*) {
  throw Error(GB_ERR_INTERNAL, "tensor-core path not built");
}
}  // namespace gb
