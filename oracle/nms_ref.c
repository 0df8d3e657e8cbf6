/* ORACLE (test infrastructure): plain-C restatement of torchvision.ops.nms, the un-vendored
 * third-party routine the reference calls at ultralytics/utils/ops.py:296
 * (dependency `torchvision>=0.9.0`, unpinned, pyproject.toml:73; 0.26.0+cu128 installed here).
 *
 * Published algorithm (torchvision/csrc/ops/cpu/nms_kernel.cpp, nms_kernel_impl<float>):
 *   areas = (x2-x1)*(y2-y1) in fp32; order = stable sort of scores, descending;
 *   walk the order; an unsuppressed box i is kept and suppresses every later unsuppressed j with
 *   inter/(area_i+area_j-inter) > thr, inter = max(0,xx2-xx1)*max(0,yy2-yy1), fp32 arithmetic, the
 *   fp32 quotient compared against the *double* threshold; returns kept indices in walk order.
 * Pinned against torchvision itself in tests/test_oracle_vs_reference.py (live, build container) and against the
 * committed reference outputs tests/golden/nms.npz in tests/test_oracle_golden.py.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/_build/liboracle.so oracle/nms_ref.c
 * (-ffp-contract=off: no FMA contraction, every op rounds to fp32 like the reference's scalar code). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void merge_sort_desc(const float *s, int32_t *idx, int32_t *tmp, int32_t n) {
    for (int32_t w = 1; w < n; w *= 2) {
        for (int32_t lo = 0; lo < n; lo += 2 * w) {
            int32_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int32_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi) tmp[o++] = (s[idx[b]] > s[idx[a]]) ? idx[b++] : idx[a++]; /* stable */
            while (a < mid) tmp[o++] = idx[a++];
            while (b < hi) tmp[o++] = idx[b++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int32_t));
    }
}

/* boxes: n x 4 (x1,y1,x2,y2) fp32 row-major; keep: capacity n. Returns number kept. */
int32_t oracle_nms(const float *boxes, const float *scores, int32_t n, double thr, int32_t *keep) {
    if (n <= 0) return 0;
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n * 2);
    float *areas = (float *)malloc(sizeof(float) * (size_t)n);
    uint8_t *sup = (uint8_t *)calloc((size_t)n, 1);
    for (int32_t i = 0; i < n; i++) {
        order[i] = i;
        areas[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    }
    merge_sort_desc(scores, order, order + n, n);
    int32_t nk = 0;
    for (int32_t _i = 0; _i < n; _i++) {
        int32_t i = order[_i];
        if (sup[i]) continue;
        keep[nk++] = i;
        float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3], ia = areas[i];
        for (int32_t _j = _i + 1; _j < n; _j++) {
            int32_t j = order[_j];
            if (sup[j]) continue;
            float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
            float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
            float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
            float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
            float w = xx2 - xx1, h = yy2 - yy1;
            w = w > 0.0f ? w : 0.0f;
            h = h > 0.0f ? h : 0.0f;
            float inter = w * h;
            float ovr = inter / (ia + areas[j] - inter);
            if ((double)ovr > thr) sup[j] = 1;
        }
    }
    free(order);
    free(areas);
    free(sup);
    return nk;
}
