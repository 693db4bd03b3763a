/* abi_smoke.c — drives libretinapost.so from plain C through include/retinapost.h: no Python, no torch.
 * Build:  gcc tests/abi_smoke.c -Iinclude -I/usr/local/cuda/include -L<dir of libretinapost.so> -lretinapost \
 *             -L/usr/local/cuda/lib64 -lcudart -lm -o abi_smoke
 * Checks: create/anchors/workspace/detect on device buffers, rpp_detect_host on host buffers gives the same
 * detections, the TPU-branch flag and the EfficientNMS entry run, error codes for a bad mode, a small workspace and
 * a Global* mode with the per-class filter. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "retinapost.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, rpp_last_error()); return 1; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static float frand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return (float)(*s >> 8) / 16777216.0f; }
static float nrand(unsigned* s) { float u = frand(s) + 1e-7f, v = frand(s); return sqrtf(-2.0f * logf(u)) * cosf(6.2831853f * v); }

int main(void) {
  static const double areas[5] = {1024.0, 4096.0, 16384.0, 65536.0, 262144.0};
  static const double ratios[3] = {0.5, 1.0, 2.0};
  static const double scales[3] = {1.0, 1.2599210498948732, 1.5874010519681994};
  rpp_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.H = cfg.W = 320; cfg.min_level = 3; cfg.max_level = 7; cfg.num_classes = 8;
  cfg.n_areas = 5; cfg.areas = areas; cfg.n_ratios = 3; cfg.aspect_ratios = ratios; cfg.n_scales = 3; cfg.scales = scales;
  cfg.box_variance[0] = cfg.box_variance[1] = 0.1f; cfg.box_variance[2] = cfg.box_variance[3] = 0.2f;
  cfg.mode = RPP_PER_CLASS_HARD_NMS; cfg.iou_threshold = 0.5f; cfg.score_threshold = 0.05f; cfg.soft_nms_sigma = 0.5f;
  cfg.pre_nms_top_k = 5000; cfg.filter_per_class = 1; cfg.max_detections = 100; cfg.soft_ignores_iou_threshold = 1;

  void* h = NULL;
  rpp_config bad = cfg;
  bad.mode = 9;
  if (rpp_create(&bad, &h) != RPP_EMODE) { fprintf(stderr, "bad mode not rejected\n"); return 1; }
  CHECK(rpp_create(&cfg, &h));
  const long N = rpp_num_anchors(h);
  if (N != 19206 || rpp_num_levels(h) != 5) { fprintf(stderr, "unexpected anchor count %ld\n", N); return 1; }
  long bounds[6];
  CHECK(rpp_anchor_boundaries(h, bounds));
  if (bounds[5] != N || bounds[1] != 14400) { fprintf(stderr, "bad boundaries\n"); return 1; }

  const int B = 3, C = 8, M = 100;
  const size_t nl = (size_t)B * N * C, nd = (size_t)B * N * 4;
  float* hl = (float*)malloc(nl * sizeof(float));
  float* hd = (float*)malloc(nd * sizeof(float));
  unsigned seed = 12345u;
  for (size_t i = 0; i < nl; ++i) hl[i] = nrand(&seed);
  for (size_t i = 0; i < nd; ++i) { float v = 0.5f * nrand(&seed); hd[i] = v > 4.f ? 4.f : (v < -4.f ? -4.f : v); }
  float *dl, *dd, *db, *ds; int *dc, *dv; void* ws;
  const size_t wsb = rpp_workspace_bytes(h, B, 0);
  CU(cudaMalloc((void**)&dl, nl * 4)); CU(cudaMalloc((void**)&dd, nd * 4));
  CU(cudaMalloc((void**)&db, (size_t)B * M * 16)); CU(cudaMalloc((void**)&ds, (size_t)B * M * 4));
  CU(cudaMalloc((void**)&dc, (size_t)B * M * 4)); CU(cudaMalloc((void**)&dv, (size_t)B * 4)); CU(cudaMalloc(&ws, wsb));
  CU(cudaMemcpy(dl, hl, nl * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dd, hd, nd * 4, cudaMemcpyHostToDevice));
  CHECK(rpp_detect(h, dd, dl, B, db, ds, dc, dv, ws, wsb, NULL));
  CU(cudaDeviceSynchronize());
  if (rpp_last_launch_count() < 4) { fprintf(stderr, "no kernels launched?\n"); return 1; }
  float sc[300], sc2[300], bx[1200], bx2[1200]; int cl[300], cl2[300], va[3], va2[3];
  CU(cudaMemcpy(sc, ds, sizeof(sc), cudaMemcpyDeviceToHost)); CU(cudaMemcpy(bx, db, sizeof(bx), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(cl, dc, sizeof(cl), cudaMemcpyDeviceToHost)); CU(cudaMemcpy(va, dv, sizeof(va), cudaMemcpyDeviceToHost));
  CHECK(rpp_detect_host(h, 0, hd, hl, B, bx2, sc2, cl2, va2));
  if (memcmp(sc, sc2, sizeof(sc)) || memcmp(bx, bx2, sizeof(bx)) || memcmp(cl, cl2, sizeof(cl)) || memcmp(va, va2, sizeof(va))) {
    fprintf(stderr, "rpp_detect_host differs from rpp_detect\n"); return 1;
  }
  for (int b = 0; b < B; ++b) {
    if (va[b] != 100) { fprintf(stderr, "image %d: %d detections\n", b, va[b]); return 1; }
    for (int i = 1; i < M; ++i) if (sc[b * M + i] > sc[b * M + i - 1]) { fprintf(stderr, "scores not sorted\n"); return 1; }
    for (int i = 0; i < M; ++i) if (cl[b * M + i] < 0 || cl[b * M + i] >= C) { fprintf(stderr, "bad class\n"); return 1; }
  }
  /* EfficientNMS_TRT-shaped entry (anchors = the handle's table): counts in range, scores sorted, classes valid */
  {
    int *dn; int nv[3];
    CU(cudaMalloc((void**)&dn, (size_t)B * 4));
    CHECK(rpp_efficient_nms(h, dd, dl, NULL, B, dn, db, ds, dc, ws, wsb, NULL));
    CU(cudaMemcpy(nv, dn, sizeof(nv), cudaMemcpyDeviceToHost)); CU(cudaMemcpy(sc2, ds, sizeof(sc2), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(cl2, dc, sizeof(cl2), cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b) {
      if (nv[b] != 100) { fprintf(stderr, "efficient_nms image %d: %d detections\n", b, nv[b]); return 1; }
      for (int i = 1; i < M; ++i) if (sc2[b * M + i] > sc2[b * M + i - 1]) { fprintf(stderr, "efficient_nms scores not sorted\n"); return 1; }
      for (int i = 0; i < M; ++i) if (cl2[b * M + i] < 0 || cl2[b * M + i] >= C) { fprintf(stderr, "efficient_nms bad class\n"); return 1; }
    }
    cudaFree(dn);
  }
  /* tpu_semantics: the TPUStrategy branch of PerClassHardNMS; a non-positive IoU threshold is rejected with it */
  {
    void* ht = NULL;
    rpp_config tc = cfg;
    tc.tpu_semantics = 1;
    CHECK(rpp_create(&tc, &ht));
    const size_t wt = rpp_workspace_bytes(ht, B, 0);
    void* wst; CU(cudaMalloc(&wst, wt));
    CHECK(rpp_detect(ht, dd, dl, B, db, ds, dc, dv, wst, wt, NULL));
    CU(cudaMemcpy(va2, dv, sizeof(va2), cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b) if (va2[b] != 100) { fprintf(stderr, "tpu branch image %d: %d detections\n", b, va2[b]); return 1; }
    cudaFree(wst);
    CHECK(rpp_destroy(ht));
    tc.iou_threshold = 0.0f;
    if (rpp_create(&tc, &ht) != RPP_EINVAL) { fprintf(stderr, "tpu_semantics with iou_threshold 0 accepted\n"); return 1; }
  }
  /* workspace too small -> RPP_EWORKSPACE; Global* + per-class filter -> RPP_ECOMBO */
  if (rpp_detect(h, dd, dl, B, db, ds, dc, dv, ws, 1024, NULL) != RPP_EWORKSPACE) { fprintf(stderr, "small workspace accepted\n"); return 1; }
  CHECK(rpp_destroy(h));
  cfg.mode = RPP_GLOBAL_SOFT_NMS;
  CHECK(rpp_create(&cfg, &h));
  if (rpp_detect(h, dd, dl, B, db, ds, dc, dv, ws, wsb, NULL) != RPP_ECOMBO) { fprintf(stderr, "Global+per-class filter accepted\n"); return 1; }
  CHECK(rpp_destroy(h));
  printf("abi_smoke ok: N=%ld, valid=[%d,%d,%d], top score %.6f\n", N, va[0], va[1], va[2], sc[0]);
  return 0;
}
