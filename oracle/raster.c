/* CPU restatement of the IUV rasteriser -- TEST INFRASTRUCTURE (see oracle/__init__.py).
 *
 * PARITY UNPINNED: the arithmetic lives in the third-party `neural_renderer` package
 * (daniilidis-group fork, reference README.md:22, requirements.txt:1), which is absent
 * from /root/reference and from this image; the reference ships no tests/golden vectors.
 * This file restates that package's published forward algorithm
 *   projection()                         (neural_renderer/projection.py)
 *   forward_face_index_map_cuda_kernel_1/2 and forward_texture_sampling_cuda_kernel
 *                                        (neural_renderer/cuda/rasterize_cuda_kernel.cu)
 *   the final vertical flip              (neural_renderer/rasterize.py)
 * anchored on the reference's call site utils/renderer.py:207-298 (IUV_Renderer):
 *   K=[[f,0,o/2],[0,f,o/2],[0,0,1]], R=I, t=[tx,ty,2f/(o*s+1e-9)], dist_coeffs=0,
 *   image_size=56, orig_size=224, near=0.1, far=100, fill_back=False,
 *   anti_aliasing=False, ambient light 1 / directional 0 (identity on textures).
 *
 * All arithmetic is IEEE fp32 without FMA contraction (build with -ffp-contract=off);
 * the CUDA kernel uses __f*_rn intrinsics in the same order, so face winners are bit-exact.
 *
 * tex_mode 0: pixel = the winning face's constant texture (SURVEY section 8c contract).
 * tex_mode 1: upstream's texture_size==1 sampling quirk: texture_index_float clamps to
 *             -eps (eps=1e-3), so the 8-corner blend reads faces f..f+3 with weights
 *             1.001^3, 3*1.001^2*(-0.001), 3*1.001*1e-6, -1e-9 (indices clamped to nf-1 here;
 *             upstream reads out of bounds for the last three faces).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void project(const float* v, const float* cam, float focal, int orig, float* out) {
    float tz = (2.0f * focal) / ((float)orig * cam[0] + 1e-9f);
    float x = v[0] + cam[1];
    float y = v[1] + cam[2];
    float z = v[2] + tz;
    float zz = z + 1e-9f;
    float x_ = x / zz;
    float y_ = y / zz;
    float c = (float)orig / 2.0f;
    float u = x_ * focal + c;
    float w = y_ * focal + c;
    w = (float)orig - w;
    u = 2.0f * (u - c) / (float)orig;
    w = 2.0f * (w - c) / (float)orig;
    out[0] = u; out[1] = w; out[2] = z;
}

/* verts [B,nv,3]; cam [B,3]; vert_mapping [nvdp] (0-based into nv); faces [nf,3] (into nvdp);
 * textures [nf,3]; img [B,3,S,S]; face_idx [B,S,S] (after the vertical flip, -1 = background);
 * depth [B,S,S] (after flip; far where empty). face_idx/depth may be NULL. */
void oracle_raster_iuv(int B, int nv, const float* verts, const float* cam,
                       int nvdp, const int32_t* vert_mapping, int nf, const int32_t* faces,
                       const float* textures, int orig, int S, float focal, float near_, float far_,
                       int tex_mode, float* img, int32_t* face_idx, float* depth) {
    float* pv = (float*)malloc(sizeof(float) * (size_t)nvdp * 3);
    float* fc = (float*)malloc(sizeof(float) * (size_t)nf * 9);
    float* finv = (float*)malloc(sizeof(float) * (size_t)nf * 9);
    unsigned char* front = (unsigned char*)malloc((size_t)nf);
    const float is = (float)S;
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < nvdp; ++i)
            project(verts + ((size_t)b * nv + vert_mapping[i]) * 3, cam + b * 3, focal, orig, pv + i * 3);
        for (int f = 0; f < nf; ++f) {
            float* face = fc + (size_t)f * 9;
            for (int k = 0; k < 3; ++k) memcpy(face + 3 * k, pv + (size_t)faces[f * 3 + k] * 3, 12);
            front[f] = !((face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0]));
            if (!front[f]) continue;
            float p[3][2];
            for (int n = 0; n < 3; ++n)
                for (int d = 0; d < 2; ++d) p[n][d] = 0.5f * (face[3 * n + d] * is + is - 1.0f);
            float fi[9] = {
                p[1][1] - p[2][1], p[2][0] - p[1][0], p[1][0] * p[2][1] - p[2][0] * p[1][1],
                p[2][1] - p[0][1], p[0][0] - p[2][0], p[2][0] * p[0][1] - p[0][0] * p[2][1],
                p[0][1] - p[1][1], p[1][0] - p[0][0], p[0][0] * p[1][1] - p[1][0] * p[0][1]};
            float den = p[2][0] * (p[0][1] - p[1][1]) + p[0][0] * (p[1][1] - p[2][1]) + p[1][0] * (p[2][1] - p[0][1]);
            for (int k = 0; k < 9; ++k) finv[(size_t)f * 9 + k] = fi[k] / den;
        }
        for (int yi = 0; yi < S; ++yi) {
            for (int xi = 0; xi < S; ++xi) {
                const float yp = (float)(2 * yi + 1 - S) / is;
                const float xp = (float)(2 * xi + 1 - S) / is;
                float depth_min = far_;
                int fmin = -1;
                for (int f = 0; f < nf; ++f) {
                    if (!front[f]) continue;
                    const float* face = fc + (size_t)f * 9;
                    if (((yp - face[1]) * (face[3] - face[0]) < (xp - face[0]) * (face[4] - face[1])) ||
                        ((yp - face[4]) * (face[6] - face[3]) < (xp - face[3]) * (face[7] - face[4])) ||
                        ((yp - face[7]) * (face[0] - face[6]) < (xp - face[6]) * (face[1] - face[7])))
                        continue;
                    const float* fi = finv + (size_t)f * 9;
                    float w[3], wsum = 0.0f;
                    for (int k = 0; k < 3; ++k) {
                        float wk = fi[3 * k + 0] * (float)xi + fi[3 * k + 1] * (float)yi + fi[3 * k + 2];
                        wk = wk > 0.0f ? wk : 0.0f;          /* min(max(w,0),1); NaN -> 0 like CUDA fmax */
                        wk = wk < 1.0f ? wk : 1.0f;
                        w[k] = wk; wsum += wk;
                    }
                    for (int k = 0; k < 3; ++k) w[k] /= wsum;
                    const float zp = 1.0f / (w[0] / face[2] + w[1] / face[5] + w[2] / face[8]);
                    if (zp <= near_ || far_ <= zp) continue;
                    if (zp < depth_min) { depth_min = zp; fmin = f; }
                }
                /* vertical flip: output row S-1-yi */
                const int yo = S - 1 - yi;
                const size_t o = ((size_t)b * S + yo) * S + xi;
                if (face_idx) face_idx[o] = fmin;
                if (depth) depth[o] = depth_min;
                for (int c = 0; c < 3; ++c) {
                    float val = 0.0f;
                    if (fmin >= 0) {
                        if (tex_mode == 0) {
                            val = textures[(size_t)fmin * 3 + c];
                        } else {
                            /* 8-corner blend of forward_texture_sampling_cuda_kernel at texture_size 1 */
                            const float tif = -1e-3f;                 /* min(max(0,0), ts-1-eps) */
                            const float fr = tif - (float)(int)tif;   /* -0.001 */
                            for (int pn = 0; pn < 8; ++pn) {
                                float ww = 1.0f; int isc = 0;
                                for (int k = 0; k < 3; ++k) {
                                    if (((pn >> k) % 2) == 0) { ww *= 1.0f - fr; }
                                    else { ww *= fr; isc += 1; }
                                }
                                int ff = fmin + isc; if (ff > nf - 1) ff = nf - 1;
                                val += ww * textures[(size_t)ff * 3 + c];
                            }
                        }
                    }
                    img[(((size_t)b * 3 + c) * S + yo) * S + xi] = val;
                }
            }
        }
    }
    free(pv); free(fc); free(finv); free(front);
}
