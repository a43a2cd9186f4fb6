// Device-side job table and kernel launchers (implemented in mobi_kernels.cu).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/mobicuda.h"

namespace mobi {

// One (stream, frame) unit of work inside a lock-step batch.  All pointers are device addresses.
struct alignas(16) DevJob {
    const mobi_frame_hdr* hdr;
    const mobi_mb* mbs;
    const mobi_part* parts;
    const mobi_op* ops;
    const mobi_coef* coefs;
    const uint32_t* intra;   // raster indices of the frame's intra MBs
    uint8_t* dst;            // luma plane of the picture being written (Stride*H), chroma follows at +Stride*H
    const uint8_t* ref[5];   // ref[k-1]: luma plane of ring picture k (Y[k], MD:413); null when absent
    uint32_t* flags;         // per-MB completion stamps of this stream (intra wavefront)
    uint32_t n_mb, n_intra;
    uint32_t dst_pic;        // index of dst / ref[k-1] along the picture axis of the ring tensor (TMA coordinates)
    uint32_t ref_pic[5];
};

// One intra macroblock to reconstruct: its descriptor travels with the work item so that the warp needs no dependent
// loads to find it.  wait: bit 0 left (m-1), 1 top-left (m-mbw-1), 2 top (m-mbw), 3 top-right (m-mbw+1): neighbours that
// are intra macroblocks of the same picture and whose pixels the predictors read -- their completion stamps are awaited.
struct alignas(16) IntraWork { uint32_t job, mb, info, first_op, first_coef, wait, pad[2]; };

struct Geom {
    int W, H, S, log2S, mbw, mbh, version;
};

// Constant tables of the kernels; once per device, before the first launch.
cudaError_t init_kernel_tables();
// The ring of pictures as TMA tensors.  A picture is ring_rows rows of Stride bytes (luma rows, chroma rows, padding), pictures
// follow each other without a gap, so the whole ring is ONE rank-2 u8 tensor:
//   l2: (Stride, rows of all pictures), box 32x17 -- a luma window;
//   c3: (Stride/2, 2, rows of all pictures) -- each row split into its U and V halves -- box 32x2x9: the U and the V window of a leaf;
//   l3 / c4: the same with the picture as a coordinate of its own (what k_inter_chunk takes).
struct InterMaps { CUtensorMap l2, c3, l3, c4; int ring_rows; };
// Inter macroblocks of every job: the fused kernel k_inter_chunk by default; MOBI_INTER_KERNEL=v3 selects k_inter_v3, =split motion
// compensation (k_mc) followed by dequantisation + inverse transforms added in place (k_res).  Persistent warps draw chunks of
// macroblocks through tickets[0] (and tickets[16] for k_res), monotonic like launch_intra's -- ticket_base[0..1] are advanced by
// what this launch draws.  between[0..1] (optional) are recorded back to back after the first kernel (per-kernel timing).
cudaError_t launch_inter(const DevJob* jobs, int n_jobs, Geom g, const InterMaps& tm, int sm_count, uint32_t* tickets, uint32_t* ticket_base,
                         cudaStream_t st, cudaEvent_t* between);
// Which inter kernel launch_inter runs (environment: MOBI_INTER_KERNEL), for reports.
const char* inter_kernel_name();
// Intra macroblocks (I-frames and intra MBs of P-frames) as a dependency wavefront.  One warp per MB; work is handed
// out through an atomic ticket in dependency order (greatest height first) so that a waiting warp's dependencies are always running.
// *warps_launched receives the number of warps started: each draws exactly one ticket past n_work, so the
// next launch's ticket_base is ticket_base + n_work + *warps_launched.
cudaError_t launch_intra(const DevJob* jobs, const IntraWork* work, uint32_t n_work, uint32_t* ticket, uint32_t ticket_base,
                         uint32_t stamp, Geom g, uint32_t max_warps, cudaStream_t st, uint32_t* warps_launched);
// I-pictures: one CTA per picture.  pics = device array of n_pics {uint32 job, uint32 work_base}; the picture's work items lie in
// raster order at work[work_base + m].
// Every CTA adds 1 to *resident when it starts (launch_gate waits on that count).
cudaError_t launch_intra_key(const DevJob* jobs, const IntraWork* work, const void* pics, int n_pics, Geom g, uint32_t* resident, cudaStream_t st);
// One tiny CTA that returns once *resident has reached target (wrap-safe) or after ~0.2 ms.
cudaError_t launch_gate(const uint32_t* resident, uint32_t target, cudaStream_t st);
// Y/UV planes of n pictures -> BGRA (MD:260-323). srcs = device array of luma plane pointers; picture i goes to
// dst + i*dst_picture_bytes with dst_pitch bytes per row.
cudaError_t launch_bgra(const uint8_t* const* srcs, int n, uint8_t* dst, int dst_pitch, size_t dst_picture_bytes, Geom g, cudaStream_t st);
// Compares the bytes of k_bgra's Moflex colour arithmetic with those of the reference's float sequence (IEEE division included)
// for every possible (Y, U, V) on the device; mismatches_host[0] receives the number of differing triples, [1] one of them.
cudaError_t selftest_bgra(unsigned long long* mismatches_host);
// Strided planes of n pictures -> tight I420 (n * W*H*3/2 bytes). srcs = device array of luma plane pointers.
cudaError_t launch_pack_i420(const uint8_t* const* srcs, int n, uint8_t* dst, Geom g, cudaStream_t st);

}  // namespace mobi
