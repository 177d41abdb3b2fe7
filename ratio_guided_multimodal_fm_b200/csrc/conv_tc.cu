// tcgen05 implicit-GEMM convolution -- placeholder until the kernel lands (fp32 mode does not use it).
#include "net.h"
namespace rgfm {
struct TcConv { int dummy; };
void TcDeleter::operator()(TcConv* p) const { delete p; }
bool tc_available() {
  int dev = 0; cudaDeviceProp p{};
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) return false;
  return p.major == 10;
}
TcPtr make_tc_conv(const TcConvDesc&, DevPool&) { throw Error("bf16 tensor-core conv not built yet"); }
void launch_conv_tc(TcConv&, const ConvArgs&, cudaStream_t) { throw Error("bf16 tensor-core conv not built yet"); }
}  // namespace rgfm
