// libpriorcorr: ABI bookkeeping — version, build info, thread-local error string.
#include <stdarg.h>
#include <stdio.h>

#include "pf_common.cuh"

namespace pf {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// Launch errors only (cudaGetLastError is non-blocking); asynchronous faults surface at the
// caller's next synchronisation, exactly like a torch op.
int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

}  // namespace pf

extern "C" {

int pf_abi_version(void) { return PF_ABI_VERSION; }

const char *pf_last_error(void) { return pf::g_error; }

#define PF_STR2(x) #x
#define PF_STR(x) PF_STR2(x)
const char *pf_build_info(void) { return "sm_100a;tcgen05;tma;abi=" PF_STR(PF_ABI_VERSION); }

}  // extern "C"
