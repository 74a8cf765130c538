/* Plain-C restatement of box_nms (TEST INFRASTRUCTURE - see oracle/__init__.py).
 *
 * Follows /root/reference/superpoint/superpoint/models/model_utils/sp_utils.py:4-28:
 *   candidates = pixels with prob >= min_prob, in row-major order            (sp_utils.py:6)
 *   boxes      = [y - s/2, x - s/2, y + s/2, x + s/2] in fp32                (sp_utils.py:8-10)
 *   keep       = torchvision.ops.nms(boxes, scores, iou)                     (sp_utils.py:14)
 *   optional top-k of the kept scores                                        (sp_utils.py:20-23)
 *   scatter kept scores into a zero map                                      (sp_utils.py:26-27)
 *
 * torchvision.ops.nms (0.15.2 pinned by the reference; 0.26 in this image) is not under
 * /root/reference; its published algorithm is restated here: stable sort by score descending, visit in
 * that order, keep a box iff it is not suppressed, a kept box suppresses every later box whose
 *   inter / (area_i + area_j - inter) > iou      (fp32 arithmetic, no +1 on widths).
 * Because boxes are equal-sized and centred on integer pixels, only neighbours with |dx|,|dy| < s can
 * overlap, so suppression is applied through a local window instead of the O(N^2) pair loop.
 * top-k tie order is unspecified in torch.topk; here (score desc, row-major index asc).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float s; int32_t idx; } cand_t;

static int cmp_desc(const void *a, const void *b) {
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx); /* stable: lower row-major index first */
}

static int suppresses(float size, float iou, int dy, int dx) {
    /* boxes centred at (0,0) and (dy,dx), fp32 as torchvision computes them */
    float h = size / 2.0f;
    float ay1 = 0.0f - h, ax1 = 0.0f - h, ay2 = 0.0f + h, ax2 = 0.0f + h;
    float by1 = (float)dy - h, bx1 = (float)dx - h, by2 = (float)dy + h, bx2 = (float)dx + h;
    float areaa = (ay2 - ay1) * (ax2 - ax1), areab = (by2 - by1) * (bx2 - bx1);
    float yy1 = fmaxf(ay1, by1), xx1 = fmaxf(ax1, bx1), yy2 = fminf(ay2, by2), xx2 = fminf(ax2, bx2);
    float w = fmaxf(0.0f, yy2 - yy1), hh = fmaxf(0.0f, xx2 - xx1);
    float inter = w * hh;
    float ovr = inter / (areaa + areab - inter);
    return ovr > iou;
}

int spn_oracle_box_nms(const float *prob, int H, int W, float size, float iou, float min_prob, int keep_top_k,
                       float *out) {
    int n = 0, i, r = (int)ceilf(size);
    cand_t *c = (cand_t *)malloc(sizeof(cand_t) * (size_t)H * W);
    uint8_t *dead = (uint8_t *)calloc((size_t)H * W, 1);
    int win = 2 * r + 1;
    uint8_t *foot = (uint8_t *)malloc((size_t)win * win);
    int nk = 0;
    if (!c || !dead || !foot) return -1;
    for (i = 0; i < H * W; ++i)
        if (prob[i] >= min_prob) { c[n].s = prob[i]; c[n].idx = i; ++n; }
    qsort(c, (size_t)n, sizeof(cand_t), cmp_desc);
    for (int dy = -r; dy <= r; ++dy)
        for (int dx = -r; dx <= r; ++dx)
            foot[(dy + r) * win + dx + r] = (uint8_t)((dy || dx) ? suppresses(size, iou, dy, dx) : 0);
    memset(out, 0, sizeof(float) * (size_t)H * W);
    for (i = 0; i < n; ++i) {
        int idx = c[i].idx, y = idx / W, x = idx % W;
        if (dead[idx]) continue;
        if (keep_top_k > 0 && nk >= keep_top_k) break; /* sorted order == top-k order (score desc, index asc) */
        out[idx] = c[i].s;
        ++nk;
        for (int dy = -r; dy <= r; ++dy) {
            int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            for (int dx = -r; dx <= r; ++dx) {
                int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                if (foot[(dy + r) * win + dx + r]) dead[yy * W + xx] = 1;
            }
        }
    }
    free(c); free(dead); free(foot);
    return nk;
}
