// blend.cu -- per-tile front-to-back alpha blending (forward) and its back-to-front adjoint.
// Replaces upstream renderCUDA forward/backward.
//
// One CTA per 16x16 tile, 4 AUTONOMOUS warps: warp w owns the 8x8 pixel square (w & 1, w >> 1), walks the tile's
// depth-ordered list on its own (no __syncthreads anywhere: ncu showed 25 % of the resident warps parked at the
// per-round barrier of the earlier CTA-staged version, waiting for the tile's slowest square) and every lane owns TWO
// of the square's pixels (same column, rows v and v + 4), evaluated together with packed FP32 (FFMA2 / FMUL2 /
// FADD2: two IEEE round-to-nearest results per instruction, element-wise identical to the scalar contract).
//
// Per group of 32 list entries (lane <-> entry): the lane's 32-byte record + 16-byte colour are already in
// registers (gathered one group ahead, straight from L2/L1: the four warps of a tile read the same lines);
// WARP-SQUARE CULLING -- each lane bounds its entry's exponent over the whole 8x8 square (exact minimum of the
// conic's quadratic form over the square) and a ballot keeps the entries that can reach alpha >= 1/255 somewhere in
// it; the survivors are compacted, in list order, into the warp's shared buffer; the next group's gathers are
// issued; then lane <-> pixel pair over the survivors.  The bound is conservative by construction (margins cover
// fp32 rounding), skipped entries are exactly the ones the per-pixel tests would skip for all 64 pixels, so images,
// final T and contributor counts stay bit-identical.  (scripts/blend_stats.py: at H0 46 % of the (square, entry)
// pairs survive and 97 % of the survivors contribute to some pixel; 8x4 rectangles keep 42 % of 1.9x as many pairs.)
#include "blend_common.cuh"

namespace dmgs {

__global__ void __launch_bounds__(BLK, FWD_MIN_BLOCKS)
blend_fwd_kernel(const __grid_constant__ BlendArgs a, const uint32_t *__restrict__ tile_order, const uint2 *__restrict__ ranges,
                 const uint32_t *__restrict__ gidx, const float4 *__restrict__ rec, const float4 *__restrict__ rgb4,
                 float *__restrict__ out_color, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
                 uint32_t *__restrict__ counter)
{
    __shared__ __align__(16) unsigned char s_cw[(BLK / 32) * CW_WARP_BYTES];  // per-warp compacted survivors

    const int lane = threadIdx.x & 31;
    const uint32_t a_cw = smem_addr(s_cw) + (uint32_t)(threadIdx.x >> 5) * CW_WARP_BYTES;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (bool first = true;; first = false) {  // persistent warp: one 8x8 square per iteration
    int tile, px0, py0;
    if (!next_square(a, tile_order, counter, first, lane, tile, px0, py0)) break;
    const int px = px0 + (lane & 7), pya = py0 + (lane >> 3), pyb = pya + 4;
    const bool in_a = px < a.W && pya < a.H, in_b = px < a.W && pyb < a.H;
    if (!__any_sync(0xffffffffu, in_a)) continue;  // the whole square lies outside the image
    const float pxf = (float)px;
    const f32x2 npx = pk2(-pxf, -pxf), npy = pk2(-(float)pya, -(float)pyb);
    const float rx0 = (float)px0, rx1 = (float)(px0 + 7), ry0 = (float)py0, ry1 = (float)(py0 + 7);
    const uint2 rng = ranges[tile];
    const int total = (int)(rng.y - rng.x);
    const uint32_t *__restrict__ list = gidx + rng.x;

    bool done_a = !in_a, done_b = !in_b;
    float Ta = 1.0f, Ca0 = 0.0f, Ca1 = 0.0f, Ca2 = 0.0f;
    float Tb = 1.0f, Cb0 = 0.0f, Cb1 = 0.0f, Cb2 = 0.0f;
    uint32_t last_a = 0, last_b = 0;

    // software pipeline: records of group k + 1 and indices of group k + 2 are in flight while group k is blended
    float4 ra = make_float4(0, 0, 0, 0), rb = ra, col = ra;
    uint32_t id_next = 0;
    if (lane < total) {
        const uint32_t g = list[lane];
        ra = ldg128(rec + 2 * (size_t)g);
        rb = ldg128(rec + 2 * (size_t)g + 1);
        col = ldg128(rgb4 + g);
    }
    if (32 + lane < total) id_next = list[32 + lane];

    for (int base = 0; base < total; base += 32) {
        // lane <-> entry: which of these 32 entries can touch the warp's square?
        const bool keep = base + lane < total && !cull_rect(ra.x, ra.y, ra.z, ra.w, rb.x, rb.z, rx0, rx1, ry0, ry1);
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        const int n = __popc(m);
        if (keep) cw_store(a_cw, __popc(m & lt_mask), ra, rb, col, (uint32_t)(base + lane + 1));
        if (base + 32 + lane < total) {
            ra = ldg128(rec + 2 * (size_t)id_next);
            rb = ldg128(rec + 2 * (size_t)id_next + 1);
            col = ldg128(rgb4 + id_next);
        }
        if (base + 64 + lane < total) id_next = list[base + 64 + lane];
        if (!n) continue;
        __syncwarp();
        // lane <-> pixel pair over the survivors, in list order
        DMGS_UNROLL(FWD_UNROLL)
        for (int t = 0; t < n; ++t) {
            const uint32_t cw = a_cw + CW_REC * (uint32_t)t;
            f32x2 power, alpha, dx, dy, G;
            alpha_two(cw, npx, npy, power, alpha, dx, dy, G);
            float p0, p1, a0, a1;
            upk2(power, p0, p1);
            upk2(alpha, a0, a1);
            const bool h0 = !done_a && p0 <= 0.0f && a0 >= 1.0f / 255.0f;
            const bool h1 = !done_b && p1 <= 0.0f && a1 >= 1.0f / 255.0f;
            if (h0 || h1) {
                const float4 c = lds128(cw + 48u);  // r, g, b, list position + 1
                if (h0) {
                    const float test_T = Ta * (1.0f - a0);
                    if (test_T < 0.0001f) {
                        done_a = true;
                    } else {
                        Ca0 = fma_(c.x * a0, Ta, Ca0);
                        Ca1 = fma_(c.y * a0, Ta, Ca1);
                        Ca2 = fma_(c.z * a0, Ta, Ca2);
                        Ta = test_T;
                        last_a = __float_as_uint(c.w);
                    }
                }
                if (h1) {
                    const float test_T = Tb * (1.0f - a1);
                    if (test_T < 0.0001f) {
                        done_b = true;
                    } else {
                        Cb0 = fma_(c.x * a1, Tb, Cb0);
                        Cb1 = fma_(c.y * a1, Tb, Cb1);
                        Cb2 = fma_(c.z * a1, Tb, Cb2);
                        Tb = test_T;
                        last_b = __float_as_uint(c.w);
                    }
                }
            }
        }
        __syncwarp();  // the buffer is rewritten by the next group
        if (__all_sync(0xffffffffu, done_a && done_b)) break;
    }
    const size_t HW = (size_t)a.H * a.W;
    if (in_a) {
        const size_t pix = (size_t)pya * a.W + px;
        final_T[pix] = Ta;
        n_contrib[pix] = last_a;
        out_color[pix] = fma_(Ta, a.bg[0], Ca0);
        out_color[HW + pix] = fma_(Ta, a.bg[1], Ca1);
        out_color[2 * HW + pix] = fma_(Ta, a.bg[2], Ca2);
    }
    if (in_b) {
        const size_t pix = (size_t)pyb * a.W + px;
        final_T[pix] = Tb;
        n_contrib[pix] = last_b;
        out_color[pix] = fma_(Tb, a.bg[0], Cb0);
        out_color[HW + pix] = fma_(Tb, a.bg[1], Cb1);
        out_color[2 * HW + pix] = fma_(Tb, a.bg[2], Cb2);
    }
    }  // next square
}

int launch_blend_fwd(const dmgs_params *prm, const void *geom, const GeomLayout &GL, const void *binning,
                     const BinLayout &BL, float *out_color, void *image, const ImgLayout &IL, cudaStream_t s)
{
    BlendArgs a;
    a.W = prm->image_width; a.H = prm->image_height;
    a.gx = (a.W + DMGS_TILE - 1) / DMGS_TILE; a.gy = (a.H + DMGS_TILE - 1) / DMGS_TILE;
    for (int i = 0; i < 3; ++i) a.bg[i] = prm->bg[i];
    if (a.W <= 0 || a.H <= 0) return 0;
    uint32_t *counter = at<uint32_t>(image, IL.counter);
    DMGS_CUDA(cudaMemsetAsync(counter, 0, sizeof(uint32_t), s));
    blend_fwd_kernel<<<blend_grid(a, 0), BLK, 0, s>>>(a, BL.has_order ? at<uint32_t>(binning, BL.tile_order) : nullptr, at<uint2>(binning, BL.ranges), at<uint32_t>(binning, BL.gidx),
                                                      at<float4>(geom, GL.rec), at<float4>(geom, GL.rgb), out_color,
                                                      at<float>(image, IL.final_T), at<uint32_t>(image, IL.n_contrib), counter);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
