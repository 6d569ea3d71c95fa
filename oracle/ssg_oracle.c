/*
 * oracle/ssg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the Self-Similarity-Graph (SSG) hot path of
 * ChrisDud0257/SSL.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load this library; the product
 * (ssl_b200/) never does.
 *
 * What it restates (paths relative to the reference checkout):
 *   GAN-Based-SR/basicsr/losses/loss_util.py:182-229      ssl_pytorch (executable spec)
 *   GAN-Based-SR/basicsr/losses/similarity/similarity.cu:5-54    raw distance, forward
 *   GAN-Based-SR/basicsr/losses/similarity/similarity.cu:73-131  raw distance, backward
 *   GAN-Based-SR/basicsr/losses/similarity/similaritywrapper.py:59-69  reflect pad + nonzero
 *   GAN-Based-SR/scripts/data_preparation/generate_mask.py:22-31  Laplacian edge mask
 *
 * Every routine exists twice (suffix _f32 / _f64) so the same loops give an
 * fp32 result in the reference's arithmetic and an fp64 "ground truth".
 * Coordinates are UNPADDED image coordinates; the reference's reflect pad by
 * P = k_s/2 (similaritywrapper.py:64-65, loss_util.py:189-191) is applied by
 * index mapping, which is the same thing as materialising the padded image.
 *
 * Parity status: pinned against the real reference `ssl_pytorch` run in the
 * build container (oracle/make_golden.py -> the .npz files under tests/golden) because the
 * reference ships no tests or golden vectors for this path.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* F.pad(mode="reflect"): padded index u -> source index (no edge repeat). */
static inline int reflect_index(int v, int n) {
    if (v < 0) v = -v;
    if (v > n - 1) v = 2 * (n - 1) - v;
    return v;
}

int ssg_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define DEFINE_ORACLE(T, SUF, EXPFN)                                                         \
                                                                                              \
/* Raw patch distance q[n, i, j]  (similarity.cu:5-54 == loss_util.py:189-223).              \
 * For edge pixel p=(py,px) and search offset (i,j) in [0,ks)^2:                             \
 *   q = sum_c sum_{a,b in [-K,K]} t^2,                                                      \
 *   t = I[c,p+(a,b)] - I[c,p-P+(i,j)+(a,b)]   if 0<=i+a<ks and 0<=j+b<ks                    \
 *   t = I[c,p+(a,b)]                          otherwise (zero-padded unfold,                \
 *                                             loss_util.py:208-209 / similarity.cu:43-47)   \
 * with I read through the reflect pad.  pos = [mc,2] (y,x), row-major order. */             \
void ssg_raw_distance_##SUF(const T* img, int C, int H, int W, const int32_t* pos, int mc,   \
                            int ks, int kw, T* q) {                                          \
    const int P = ks / 2, K = kw / 2;                                                        \
    _Pragma("omp parallel for schedule(static)")                                             \
    for (int n = 0; n < mc; ++n) {                                                           \
        const int py = pos[2 * n], px = pos[2 * n + 1];                                      \
        for (int i = 0; i < ks; ++i)                                                         \
            for (int j = 0; j < ks; ++j) {                                                   \
                T acc = 0;                                                                   \
                for (int c = 0; c < C; ++c) {                                                \
                    const T* plane = img + (size_t)c * H * W;                                \
                    for (int a = -K; a <= K; ++a)                                            \
                        for (int b = -K; b <= K; ++b) {                                      \
                            const T ctr = plane[reflect_index(py + a, H) * W +               \
                                                reflect_index(px + b, W)];                   \
                            T t = ctr;                                                       \
                            if (i + a >= 0 && i + a < ks && j + b >= 0 && j + b < ks)        \
                                t = ctr - plane[reflect_index(py - P + i + a, H) * W +       \
                                                reflect_index(px - P + j + b, W)];           \
                            acc += t * t;                                                    \
                        }                                                                    \
                }                                                                            \
                q[((size_t)n * ks + i) * ks + j] = acc;                                      \
            }                                                                                \
    }                                                                                        \
}                                                                                            \
                                                                                              \
/* Tail of ssl_pytorch / ssl_cuda (loss_util.py:224-228 == :234-243):                        \
 *   d = q / (C*kw^2);  e = exp(-1*d/sigma);  s = (1/(sum_j e_j + eps)) * e  if normalise */ \
void ssg_rows_from_distance_##SUF(const T* q, int mc, int ks, int kw, int C, T sigma,        \
                                  int normalise, T eps, T* rows) {                           \
    const int L = ks * ks;                                                                   \
    const T denom = (T)C * (T)(kw * kw);                                                     \
    _Pragma("omp parallel for schedule(static)")                                             \
    for (int n = 0; n < mc; ++n) {                                                           \
        const T* qr = q + (size_t)n * L;                                                     \
        T* sr = rows + (size_t)n * L;                                                        \
        T z = 0;                                                                             \
        for (int j = 0; j < L; ++j) {                                                        \
            const T d = qr[j] / denom;                                                       \
            sr[j] = EXPFN((T)-1 * d / sigma);                                                \
            z += sr[j];                                                                      \
        }                                                                                    \
        if (normalise) {                                                                     \
            const T r = (T)1 / (z + eps);                                                    \
            for (int j = 0; j < L; ++j) sr[j] = r * sr[j];                                   \
        }                                                                                    \
    }                                                                                        \
}                                                                                            \
                                                                                              \
/* Adjoint of the tail: given rows s (as produced above) and g_s = dL/ds,                    \
 * return g_q = dL/dq  (SURVEY.md 8a-7 chain, checked against reference autograd):           \
 *   normalised:  g_e = (g_s - sum_m g_m s_m) / Z',  e = s*Z'  =>                            \
 *                g_q = -s * (g_s - dot) / (sigma*C*kw^2)                                    \
 *   otherwise :  g_q = -e * g_e / (sigma*C*kw^2)             (rows hold e)        */        \
void ssg_rows_backward_##SUF(const T* rows, const T* grows, int mc, int ks, int kw, int C,   \
                             T sigma, int normalise, T* gq) {                                \
    const int L = ks * ks;                                                                   \
    const T scale = (T)-1 / (sigma * (T)C * (T)(kw * kw));                                   \
    _Pragma("omp parallel for schedule(static)")                                             \
    for (int n = 0; n < mc; ++n) {                                                           \
        const T* s = rows + (size_t)n * L;                                                   \
        const T* g = grows + (size_t)n * L;                                                  \
        T dot = 0;                                                                           \
        if (normalise)                                                                       \
            for (int j = 0; j < L; ++j) dot += g[j] * s[j];                                  \
        for (int j = 0; j < L; ++j) gq[(size_t)n * L + j] = scale * s[j] * (g[j] - dot);     \
    }                                                                                        \
}                                                                                            \
                                                                                              \
/* Backward of the raw distance (similarity.cu:73-131) followed by the adjoint of the        \
 * reflect pad that autograd applies in the reference (similaritywrapper.py:64): every       \
 * contribution lands on the source pixel its padded index mirrors to.                       \
 * grad_img [C,H,W] is ACCUMULATED into (caller zeroes it, similaritywrapper.py:47). */      \
void ssg_raw_distance_backward_##SUF(const T* img, const T* gq, int C, int H, int W,         \
                                     const int32_t* pos, int mc, int ks, int kw,             \
                                     T* grad_img) {                                          \
    const int P = ks / 2, K = kw / 2;                                                        \
    const size_t npx = (size_t)C * H * W;                                                    \
    const int nthreads = ssg_oracle_num_threads();                                           \
    T* scratch = (T*)calloc(npx * (size_t)nthreads, sizeof(T));                              \
    _Pragma("omp parallel")                                                                  \
    {                                                                                        \
        const int tid = ssg_oracle_thread_id();                                              \
        T* G = scratch + npx * (size_t)tid;                                                  \
        _Pragma("omp for schedule(static)")                                                  \
        for (int n = 0; n < mc; ++n) {                                                       \
            const int py = pos[2 * n], px = pos[2 * n + 1];                                  \
            for (int i = 0; i < ks; ++i)                                                     \
                for (int j = 0; j < ks; ++j) {                                               \
                    const T g = gq[((size_t)n * ks + i) * ks + j];                           \
                    for (int c = 0; c < C; ++c) {                                            \
                        const T* plane = img + (size_t)c * H * W;                            \
                        T* gplane = G + (size_t)c * H * W;                                   \
                        for (int a = -K; a <= K; ++a)                                        \
                            for (int b = -K; b <= K; ++b) {                                  \
                                const int ci = reflect_index(py + a, H) * W +                \
                                               reflect_index(px + b, W);                     \
                                if (i + a >= 0 && i + a < ks && j + b >= 0 && j + b < ks) {  \
                                    const int ni = reflect_index(py - P + i + a, H) * W +    \
                                                   reflect_index(px - P + j + b, W);         \
                                    const T v = (T)2 * (plane[ci] - plane[ni]) * g;          \
                                    gplane[ci] += v;                                         \
                                    gplane[ni] -= v;                                         \
                                } else {                                                     \
                                    gplane[ci] += (T)2 * plane[ci] * g;                      \
                                }                                                            \
                            }                                                                \
                    }                                                                        \
                }                                                                            \
        }                                                                                    \
    }                                                                                        \
    for (int t = 0; t < nthreads; ++t)                                                       \
        for (size_t k = 0; k < npx; ++k) grad_img[k] += scratch[npx * (size_t)t + k];        \
    free(scratch);                                                                           \
}

static inline int ssg_oracle_thread_id(void) {
#ifdef _OPENMP
    return omp_get_thread_num();
#else
    return 0;
#endif
}

DEFINE_ORACLE(float, f32, expf)
DEFINE_ORACLE(double, f64, exp)

/*
 * Offline edge mask of the reference (generate_mask.py:22-31), restated:
 *   L    = PIL convert("L"): (19595 R + 38470 G + 7471 B + 32768) >> 16   (ITU-R 601)
 *   lap  = cv2.Laplacian(L, CV_8U): 4-neighbour kernel [[0,1,0],[1,-4,1],[0,1,0]],
 *          BORDER_REFLECT_101, saturated to [0,255]
 *   mask = lap > threshold
 * rgb is uint8 [3,H,W] (planar).  mask is float 0/1 [H,W], the form the datasets hand to
 * the loss (my_realesrgan_image_mask_dataset.py:79-86).
 */
void ssg_laplacian_mask_u8(const uint8_t* rgb, int H, int W, float threshold, float* mask) {
    uint8_t* L = (uint8_t*)malloc((size_t)H * W);
    for (int k = 0; k < H * W; ++k)
        L[k] = (uint8_t)((19595u * rgb[k] + 38470u * rgb[(size_t)H * W + k] +
                          7471u * rgb[(size_t)2 * H * W + k] + 32768u) >> 16);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            /* BORDER_REFLECT_101 == the same no-repeat mirror as reflect_index */
            const int yu = reflect_index(y - 1, H), yd = reflect_index(y + 1, H);
            const int xl = reflect_index(x - 1, W), xr = reflect_index(x + 1, W);
            int v = (int)L[yu * W + x] + L[yd * W + x] + L[y * W + xl] + L[y * W + xr] -
                    4 * (int)L[y * W + x];
            if (v < 0) v = 0;
            if (v > 255) v = 255;
            mask[y * W + x] = ((float)v > threshold) ? 1.0f : 0.0f;
        }
    free(L);
}
