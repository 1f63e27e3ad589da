/* A plain-C consumer of include/v100.h: what a cgo / JNI / FFI binding of the reference's side would compile.
 * Built with `gcc -std=c99 -Wall -Werror -pedantic` by tests/test_cabi.py (no CUDA headers, no C++), linked against
 * libv100.so, and run on the CPU-only box: it takes the address of every entry point the header declares (so a
 * signature the header and the library disagree on fails to link or warns), checks the ABI version and makes ONE
 * call that must fail loudly without a GPU instead of falling back to the host. */
#include <stdio.h>
#include <string.h>

#include "v100.h"

int main(void) {
  typedef void (*fn_t)(void);
  const fn_t table[] = {
      (fn_t)v100_abi_version, (fn_t)v100_last_error, (fn_t)v100_conv1x1,
      (fn_t)v100_conv1x1_f32out, (fn_t)v100_dwconv1d, (fn_t)v100_convtranspose1d_k5s2,
      (fn_t)v100_logmel, (fn_t)v100_ctc_finalize, (fn_t)v100_ctc_collapse,
      (fn_t)v100_ctc_best_path, (fn_t)v100_world_finalize, (fn_t)v100_embedding_ncw16,
      (fn_t)v100_lstm_layer, (fn_t)v100_lstm_workspace_bytes, (fn_t)v100_layernorm_gelu,
  };
  size_t i;
  for (i = 0; i < sizeof(table) / sizeof(table[0]); ++i)
    if (table[i] == NULL) return 2;
  if (v100_abi_version() != V100_ABI_VERSION) {
    fprintf(stderr, "ABI %d, header %d\n", v100_abi_version(), V100_ABI_VERSION);
    return 3;
  }
  /* null pointers: the argument check answers before any CUDA call, with a message */
  {
    int rc = v100_conv1x1(NULL, 0, NULL, NULL, NULL, NULL, NULL, 0, 1, 8, 8, 8, V100_ACT_NONE, V100_DTYPE_BF16, NULL);
    const char* msg = v100_last_error();
    if (rc == 0 || msg == NULL || strlen(msg) == 0) return 4;
    printf("v100 ABI %d; conv1x1(NULL...) -> %d (%s)\n", v100_abi_version(), rc, msg);
  }
  return 0;
}
