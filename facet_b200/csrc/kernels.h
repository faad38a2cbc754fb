// Internal launcher declarations (C++); the public C ABI is include/facet_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fb {

int tech_rows_per_unit(int n, int H, int W, int sms);
int launch_tech_stats(const uint8_t* d_images, int n, int H, int W, long long image_stride, int rgb_order,
                      unsigned int* d_hist256, unsigned int* d_hs_hist, long long* d_sums, int force_generic,
                      cudaStream_t stream);
int launch_gray_hsv(const uint8_t* d_image, int H, int W, int rgb_order, uint8_t* d_gray, uint8_t* d_hsv,
                    cudaStream_t stream);
int launch_hs_derive(const unsigned int* d_hs_hist, int n, double* d_out, cudaStream_t stream);

int launch_hamming_pairs(const unsigned long long* d_hashes, long long n, int max_distance, int part,
                         int nparts, int* d_pairs, long long cap, unsigned long long* d_count,
                         cudaStream_t stream);
int launch_burst_links(const unsigned long long* d_hashes, const long long* d_time_s, const unsigned char* d_flags,
                       const int* d_lo, long long n, int thr, long long window_s, double rapid_s,
                       int* d_last_slow, int* d_rapid_pairs, long long rapid_cap,
                       unsigned long long* d_rapid_count, cudaStream_t stream);
int launch_clip_preprocess(const uint8_t* d_images, int n, int H, int W, long long image_stride, int rgb_order,
                           int out_size, const int* d_hbounds, const int* d_hcoef, int hk, int h_byte_lo,
                           int h_byte_hi, const int* d_vbounds, const int* d_vcoef, int vk, int row0, int rows,
                           const float* mean3, const float* std3, uint8_t* d_tmp, float* d_out,
                           cudaStream_t stream);
int launch_roi_laplacian(const uint8_t* d_image, int H, int W, int rgb_order, const int* d_boxes, int k,
                         long long* d_out, cudaStream_t stream);

}  // namespace fb
