/* Compiled as plain C99 by tests/test_abi.py: the header must be valid C, and the entry points
 * that need no device must behave as documented when called from C. */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "perceive_cuda.h"

int main(void) {
  if (pcv_abi_version() != PCV_ABI_VERSION) return 1;
  /* embedding codec: search.rs:281-294 (little-endian f32, no header) */
  const float v[3] = {1.0f, -2.5f, 3.25f};
  uint8_t blob[12];
  if (pcv_encode_embedding(v, 3, blob, sizeof blob) != PCV_OK) return 2;
  const uint8_t want0[4] = {0x00, 0x00, 0x80, 0x3f};
  if (memcmp(blob, want0, 4) != 0) return 3;
  float back[3];
  size_t dim = 0;
  if (pcv_decode_embedding(blob, 12, back, 3, &dim) != PCV_OK || dim != 3 || memcmp(back, v, 12) != 0) return 4;
  if (pcv_decode_embedding(blob, 11, back, 3, &dim) != PCV_ERR_INVALID) return 5; /* reference panics here */
  if (strlen(pcv_last_error()) == 0) return 6;
  /* reference distance: max(0, 1 - dot/len), search.rs:274-277 */
  if (pcv_distance_from_dot(96.0f, 384) != 0.75f) return 7;
  if (pcv_distance_from_dot(500.0f, 384) != 0.0f) return 8;
  /* argument validation happens before any device work */
  pcv_index* ix = (pcv_index*)1;
  if (pcv_index_create(0, 0, PCV_F32, PCV_METRIC_DOT_REF, 0, &ix) != PCV_ERR_INVALID || ix != NULL) return 9;
  if (pcv_index_create(0, 384, (pcv_dtype)7, PCV_METRIC_DOT_REF, 0, &ix) != PCV_ERR_INVALID) return 10;
  if (pcv_index_create(0, 1024, PCV_F32_SPLIT, PCV_METRIC_DOT_REF, 0, &ix) != PCV_ERR_UNSUPPORTED) return 11;
  if (pcv_search(NULL, v, 1, 1, NULL, 0, NULL, NULL, NULL, NULL) != PCV_ERR_INVALID) return 12;
  /* one handle over several GPUs: the device list is validated before any device is touched */
  {
    const int32_t twice[2] = {0, 0};
    ix = (pcv_index*)1;
    if (pcv_index_create_multi(NULL, 2, 384, PCV_F32, PCV_METRIC_DOT_REF, 0, &ix) != PCV_ERR_INVALID || ix != NULL) return 15;
    if (pcv_index_create_multi(twice, 0, 384, PCV_F32, PCV_METRIC_DOT_REF, 0, &ix) != PCV_ERR_INVALID) return 16;
    if (pcv_index_create_multi(twice, 2, 384, PCV_F32, PCV_METRIC_DOT_REF, 0, &ix) != PCV_ERR_INVALID) return 17; /* same GPU twice */
    if (pcv_index_create_multi(twice, 17, 384, PCV_F32, PCV_METRIC_DOT_REF, 0, &ix) != PCV_ERR_INVALID) return 18;
  }
  float q[4];
  if (pcv_synthetic_rows_host(2, PCV_DIST_UNIT_SPHERE, 0, 1, 4, q) != PCV_OK) return 13;
  if (fabsf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3] - 1.0f) > 1e-5f) return 14;
  puts("abi ok");
  return 0;
}
