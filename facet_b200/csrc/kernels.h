// Internal launcher declarations (C++); the public C ABI is include/facet_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FB_GEMM_BIAS_BF16 0
#define FB_GEMM_BIAS_GELU_BF16 1
#define FB_GEMM_BIAS_RESIDUAL_F32 2
#define FB_GEMM_F32 3
#define FB_GEMM_THRESHOLD_PAIRS 4
#define FB_GEMM_F16_FLAG 16   /* OR into the mode: operands and 16-bit outputs are fp16 instead of bf16 */

namespace fb {

int tech_rows_per_unit(int n, int H, int W, int sms);
int launch_tech_stats(const uint8_t* d_images, int n, int H, int W, long long image_stride, int rgb_order,
                      unsigned int* d_hist256, unsigned int* d_hs_hist, long long* d_sums, int force_generic,
                      uint8_t* d_luma, uint8_t* d_box4, const unsigned int* box_mult4, cudaStream_t stream);
int launch_box_reduce(const uint8_t* d_images, int n, int H, int W, long long image_stride, int fx, int fy, int red_h, int red_w,
                      const unsigned int* mult4, uint8_t* d_reduced, cudaStream_t stream);
int launch_gray_hsv(const uint8_t* d_image, int H, int W, int rgb_order, uint8_t* d_gray, uint8_t* d_hsv,
                    cudaStream_t stream);
int launch_gray_plane(const uint8_t* d_image, int H, int W, int rgb_order, uint8_t* d_gray, unsigned int* d_hist256,
                      cudaStream_t stream);
size_t canny_workspace_bytes(int H, int W);
int launch_canny(const uint8_t* d_gray, int H, int W, int blur, int low, int high, void* d_ws, size_t ws_bytes,
                 uint8_t* d_edges, unsigned long long* d_count, cudaStream_t stream);
int launch_hs_derive(const unsigned int* d_hs_hist, int n, double* d_out, cudaStream_t stream);

int launch_hamming_pairs(const unsigned long long* d_hashes, long long n, int max_distance, int part,
                         int nparts, int* d_pairs, long long cap, unsigned long long* d_count,
                         cudaStream_t stream);
int launch_burst_links(const unsigned long long* d_hashes, const long long* d_time_s, const unsigned char* d_flags,
                       const int* d_lo, long long n, int thr, long long window_s, double rapid_s,
                       int* d_last_slow, int* d_rapid_pairs, long long rapid_cap,
                       unsigned long long* d_rapid_count, cudaStream_t stream);
int launch_clip_preprocess(const uint8_t* d_images, int n, int H, int W, long long image_stride, int rgb_order,
                           int out_size, const int* d_hp0, const int* d_hcpad, int hgroups, int h_px_lo,
                           int h_span_px, const int* d_vbounds, const int* d_vcoef, int vk, int row0, int rows,
                           const float* mean3, const float* std3, uint8_t* d_tmp, float* d_out,
                           const int8_t* d_tc_coef, int tc_kw, int tc_limbs, const int* d_tc_kb0, cudaStream_t stream);
int launch_resample_h_tc(const uint8_t* d_images, int n, int H, int W, long long image_stride, int out_size, const int8_t* d_coef,
                         int kw, int limbs, const int* d_kb0, int row0, int rows, uint8_t* d_tmp, int channels,
                         cudaStream_t stream);
int launch_phash(const uint8_t* d_images, int n, int H, int W, long long image_stride, int rgb_order, const int* d_hbounds,
                 const int* d_hcoef, int hk, const int* d_vbounds, const int* d_vcoef, int vk, uint8_t* d_tmp,
                 unsigned long long* d_hashes, uint8_t* d_small, double* d_dct, uint8_t* d_luma, int luma_ready,
                 const int8_t* d_tc_coef, int tc_kw, int tc_limbs, const int* d_tc_kb0, cudaStream_t stream);
int launch_orient(const uint8_t* d_src, int n, int H, int W, long long src_stride, int swap, int flip_x, int flip_y,
                  int swap_rb, uint8_t* d_dst, long long dst_stride, cudaStream_t stream);
int launch_thumbnail(const uint8_t* d_images, int n, int H, int W, long long image_stride, int fx, int fy, int red_h, int red_w,
                     const unsigned int* mult4, const int* d_hbounds, const int* d_hcoef, int hk, const int* d_vbounds,
                     const int* d_vcoef, int vk, int out_h, int out_w, int swap_rb, uint8_t* d_reduced, uint8_t* d_tmp,
                     uint8_t* d_out, int reduced_ready, cudaStream_t stream);
int launch_roi_laplacian(const uint8_t* d_image, int H, int W, int rgb_order, const int* d_boxes, int k,
                         long long* d_out, cudaStream_t stream);

int launch_gemm_bf16(const void* d_a, long long lda, const void* d_b, long long ldb, int M, int N, int K, int mode,
                     const float* d_bias, void* d_out, long long ldo, const float* d_residual, long long ldr,
                     cudaStream_t stream);
// LayerNorm fold of the ViT tower (csrc/gemm.cu GemmArgs): producer (residual epilogue) writes out16 + row_stats, consumer
// (bias / GELU epilogue) reads row_stats + ln_s.
struct GemmLnFold {
    void* out16;
    long long ldo16;
    float* row_stats;
    int ln_slots;
    const float* ln_s;
    int ln_width;
};
int launch_gemm_bf16_ln(const void* d_a, long long lda, const void* d_b, long long ldb, int M, int N, int K, int mode,
                        const float* d_bias, void* d_out, long long ldo, const float* d_residual, long long ldr,
                        const GemmLnFold* ln, cudaStream_t stream);
int launch_cosine_candidates(const void* d_emb_bf16, long long ld, int n, int row_offset, int m, int k, float tau,
                             int* d_pairs, float* d_sims, long long cap, unsigned long long* d_count,
                             cudaStream_t stream);
int launch_cosine_block(const void* d_a_bf16, int m, int a_offset, const void* d_b_bf16, int n, int b_offset, long long ld, int k, float tau,
                        int triangle, int* d_pairs, float* d_sims, long long cap, unsigned long long* d_count, cudaStream_t stream);
int launch_cosine_recheck(const float* d_emb_f32, long long ld, int k, const int* d_cand, const unsigned long long* d_ncand,
                          long long cand_cap, float tau, int* d_pairs, float* d_sims, long long cap,
                          unsigned long long* d_count, cudaStream_t stream);
int launch_f32_to_bf16(const float* d_in, void* d_out, long long n, cudaStream_t stream);
int launch_im2col_patch14(const float* d_x, int batch, void* d_out, int f16, cudaStream_t stream);
int launch_layernorm(const float* d_in, long long ld_in, int rows, const float* gamma, const float* beta,
                     const float* cls, const float* pos, void* d_out, long long ld_out, int out_bf16,
                     cudaStream_t stream);   // out_bf16: 0 = fp32, 1 = bf16, 2 = fp16
// ln_pre with the class-token / positional-embedding assembly, additionally writing the 16-bit copy of the residual stream and
// the row sums the LayerNorm fold of the first block needs (slot 0 carries the whole row, the other slots are zero)
int launch_layernorm_pre_fold(const float* d_in, long long ld_in, int rows, const float* gamma, const float* beta, const float* cls,
                              const float* pos, float* d_out, long long ld_out, void* d_out16, int f16, float* d_row_stats, int ln_slots,
                              cudaStream_t stream);
int launch_attention_tc(const void* d_qkv, int batch, void* d_out, int f16, cudaStream_t stream);   // tcgen05
int launch_vit_tail(const float* d_x, int batch, const float* g, const float* be, const float* proj, const float* w1,
                    const float* b1, const float* w2, const float* b2, const float* tags, int ntags, float* feat,
                    float* emb, float* raw, float* sims, cudaStream_t stream);
int launch_embedding_heads(const float* d_x, int n, const float* w1, const float* b1, const float* w2, const float* b2,
                           const float* tags, int ntags, float* raw, float* sims, cudaStream_t stream);
size_t jpeg_workspace_bytes(int n, int width, int height, int ncomp, int hs0, int vs0, int restart_interval, long long max_scan_bytes);
int launch_jpeg_decode(const uint8_t* d_bytes, const long long* d_scan_off, const long long* d_scan_len, const int* d_table_slot,
                       const void* d_tables, int n, int width, int height, int ncomp, const int* hs, const int* vs, const int* tq,
                       const int* td, const int* ta, int restart_interval, long long max_scan_bytes, int bgr, void* d_workspace,
                       size_t workspace_bytes, uint8_t* d_out, long long out_stride, int* d_status, cudaStream_t stream);
size_t jpeg_encode_workspace_bytes(int n, int H, int W);
size_t jpeg_encode_out_stride(int H, int W, int header_len);
int launch_jpeg_encode(const uint8_t* d_rgb, int n, int H, int W, long long image_stride, const void* d_tables, const uint8_t* d_header,
                       int header_len, void* d_ws, size_t ws_bytes, uint8_t* d_out, long long out_stride, unsigned int* d_length,
                       cudaStream_t stream);
void count_launch(int k);

// Optional per-category CUDA-event timing of the library's launches (bench.py roofline): when enabled every
// launch wrapper records an event pair on the launching stream; fb_profile_read() turns them into milliseconds.
enum ProfCat { PROF_TECH = 0, PROF_DERIVE, PROF_PREPROCESS, PROF_IM2COL, PROF_GEMM, PROF_LAYERNORM, PROF_ATTENTION,
               PROF_TAIL, PROF_COSINE, PROF_HAMMING, PROF_OTHER, PROF_NCAT };
struct ProfScope {
    ProfScope(int cat, cudaStream_t st);
    ~ProfScope();
    int slot;
    cudaStream_t stream;
    cudaEvent_t begin, end;
};
size_t vit_workspace_bytes(int batch);

}  // namespace fb
