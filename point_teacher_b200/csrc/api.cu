// C-ABI plumbing shared by every entry point: error text, launch checks, version.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace ptb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return PT_ERR_CUDA;
  }
  return PT_OK;
}

}  // namespace ptb

extern "C" const char* pt_last_error(void) { return ptb::g_err; }
extern "C" int pt_abi_version(void) { return 1; }
extern "C" const char* pt_build_arch(void) { return "sm_100a"; }
