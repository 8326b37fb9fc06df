/* Plain-C99 consumer of include/tokenpacker_b200.h, linked against libtokenpacker_b200.so: the boundary really is a C ABI
 * (no C++ types, no torch types), every declared entry point resolves at link time, and the pure size queries run without a GPU.
 * Built and run by tests/test_abi.py::test_plain_c_consumer. */
#include <stdio.h>
#include <string.h>

#include "tokenpacker_b200.h"

int main(void) {
  /* taking the address of every entry point makes the link fail if one is declared but not exported */
  const void* entry[] = {
      (const void*)&tp_strerror,          (const void*)&tp_abi_version,       (const void*)&tp_last_cuda_error,
      (const void*)&tp_packed_bytes,      (const void*)&tp_pack_weights,      (const void*)&tp_workspace_bytes,
      (const void*)&tp_forward,           (const void*)&tp_forward_layers,    (const void*)&tp_forward_allgather,
      (const void*)&tp_forward_host,      (const void*)&tp_train_saved_bytes, (const void*)&tp_backward_workspace_bytes,
      (const void*)&tp_forward_train,     (const void*)&tp_backward,          (const void*)&tp_gemm_bf16,
      (const void*)&tp_gemm_tn_bf16,      (const void*)&tp_hd_grid,           (const void*)&tp_hd_fit,
      (const void*)&tp_hd_tile,           (const void*)&tp_hd_plan,           (const void*)&tp_hd_scatter_crops,
      (const void*)&tp_gather_rows,       (const void*)&tp_hd_fill_separators, (const void*)&tp_forward_packed,
      (const void*)&tp_launch_count,      (const void*)&tp_hd_tile_batch_plan, (const void*)&tp_hd_tile_batch,
      (const void*)&tp_gemm_nn_bf16,      (const void*)&tp_pack_weights_train};
  size_t i;
  for (i = 0; i < sizeof(entry) / sizeof(entry[0]); ++i)
    if (entry[i] == NULL) return 2;
  if (tp_abi_version() != 2) return 3;
  if (strcmp(tp_strerror(TP_ERR_BAD_SCALE_FACTOR), "scale_factor must be divisible by grid size") != 0) return 4;
  if (tp_packed_bytes(4096) == 0 || tp_packed_bytes(4097) != 0) return 5;
  if (tp_workspace_bytes(64, 2, 4096) == 0 || tp_workspace_bytes(64, 5, 4096) != 0) return 6;
  {
    /* the grid selector is host arithmetic (patch_divide.py:96-105): 1088 x 1088 with patch_num 9 -> 3 x 3 */
    int hb = 0, wb = 0;
    if (tp_hd_grid(1088, 1088, 9, 336, &hb, &wb) != TP_OK || hb != 3 || wb != 3) return 7;
    if (tp_hd_grid(1088, 1088, 10, 336, &hb, &wb) != TP_ERR_BAD_PATCH_NUM) return 8;
  }
  {
    /* tiling sizes (train.py:701-708, :719-726): 1088 x 1088 on a 3 x 3 grid fills the 1008 canvas; thumbnail 336 x 336 */
    int hr = 0, wr = 0, ht = 0, wt = 0;
    if (tp_hd_fit(1088, 1088, 3, 3, &hr, &wr, &ht, &wt) != TP_OK || hr != 1008 || wr != 1008 || ht != 336 || wt != 336) return 11;
    if (tp_hd_fit(500, 700, 2, 3, &hr, &wr, &ht, &wt) != TP_OK || hr != 672 || wr != 941 || ht != 240 || wt != 336) return 12;
  }
  {
    /* slice assembly plan (llava_arch.py:139-155): a 3 x 3 image at 144 tokens per crop is 1450 rows, a 1 x 1 image 145 */
    const int hb[2] = {3, 1}, wb[2] = {3, 1};
    int64_t n_crops = 0, n_sep = 0, n_ret = 0;
    int64_t seg[11], sep[6], ret[5], cu[3];
    if (tp_hd_plan(hb, wb, 2, 144, NULL, NULL, NULL, NULL, &n_crops, &n_sep, &n_ret) != TP_OK) return 13;
    if (n_crops != 11 || n_sep != 6 || n_ret != 5) return 14;
    if (tp_hd_plan(hb, wb, 2, 144, seg, sep, ret, cu, &n_crops, &n_sep, &n_ret) != TP_OK) return 15;
    if (cu[0] != 0 || cu[1] != 1450 || cu[2] != 1595) return 16;
    if (seg[0] != 0 || seg[1] != 145 || seg[9] != 1305 || seg[10] != 1450) return 17;   /* thumbnail after the grid; next image */
  }
  {
    /* argument validation happens before any CUDA call */
    if (tp_forward(NULL, NULL, NULL, 1, 0, 0, 5, 4096, NULL, NULL, NULL, 0, NULL) != TP_ERR_BAD_SCALE_FACTOR) return 9;
    if (tp_forward(NULL, NULL, NULL, 1, 0, 0, 2, 4096, NULL, NULL, NULL, 0, NULL) != TP_ERR_INVALID_ARGUMENT) return 10;
  }
  printf("abi ok: %u entry points\n", (unsigned)(sizeof(entry) / sizeof(entry[0])));
  return 0;
}
