// libmobicuda.so runtime: the C ABI of include/mobicuda.h over the host parser (mobi_parse.cpp) and the
// sm_100a kernels (mobi_kernels.cu).  "MD:n" = LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs:n.
//
// Memory plan (per GPU):
//   ring      N streams x 6 pictures x (Stride*H luma + Stride*H/2 chroma + pad), zeroed once.  The kernels
//             only ever write pixels inside the visible W x H area, so the stride padding stays zero for the
//             life of the batch exactly as in the reference's freshly allocated planes (MD:107-108).
//   flags     N x n_MB completion stamps for the intra wavefront (stamp = launch serial, never cleared).
//   arena     one pinned host + one device buffer per in-flight step holding
//             [DevJob table | IntraWork list | per stream: hdr, mbs, parts, ops, coefs, intra list],
//             uploaded with a single cudaMemcpyAsync.  Two arenas alternate so that the host parses step
//             k+1 while the GPU reconstructs step k.
//   staged    device-resident copies of whole steps for the kernel-only replay (bench "value" leg, ncu).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "mobi_kernels.h"
#include "mobi_parse.h"

namespace mobi {
namespace {

constexpr int RING = 6;  // Y[0..5] (MD:19)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- tiny persistent thread pool: parallel_for over stream indices --------------------------------
class Pool {
public:
    explicit Pool(int n) {
        for (int i = 0; i < n; i++) th_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void run(int n, const std::function<void(int)>& fn) {
        if (th_.empty() || n <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
        {
            std::lock_guard<std::mutex> l(m_);
            fn_ = &fn; n_ = n; next_.store(0); active_ = (int)th_.size(); gen_++; failed_.store(false);
        }
        cv_.notify_all();
        work();  // the caller helps
        {
            std::unique_lock<std::mutex> l(m_);
            done_.wait(l, [this] { return active_ == 0; });
            fn_ = nullptr;
        }
        if (failed_.load()) throw std::bad_alloc();   // (the only thing the work items can throw; the C ABI maps it to MOBI_ERR_NOMEM)
    }
private:
    // No exception leaves a worker thread (std::terminate) or unwinds run() while workers still hold fn_: the first
    // failure is remembered and rethrown by run() on the calling thread once every worker is done.
    void work() {
        for (;;) {
            int i = next_.fetch_add(1);
            if (i >= n_) break;
            try { (*fn_)(i); } catch (...) { failed_.store(true); }
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            work();
            {
                std::lock_guard<std::mutex> l(m_);
                if (--active_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    std::atomic<int> next_{0};
    std::atomic<bool> failed_{false};
    int n_ = 0, active_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

struct Arena {
    uint8_t* h = nullptr;  // pinned
    uint8_t* d = nullptr;
    size_t cap = 0;
    cudaEvent_t consumed = nullptr;  // recorded after the kernels that read d
    bool pending = false;
};

// Where one step's arrays live on the device, plus the host-side facts needed to launch it again.
struct StepLayout {
    size_t bytes = 0;
    size_t jobs_off = 0, work_off = 0, key_off = 0, pics_off = 0;   // job table, intra work list, {job, work_base} of the I-pictures, luma plane of each job's new picture
    int n_jobs = 0, n_inter_jobs = 0, n_key_jobs = 0;
    uint32_t n_work = 0;      // all intra macroblocks; the list holds those of P-pictures first, then those of I-pictures
    uint32_t n_work_p = 0;    // intra macroblocks inside P-pictures
    std::vector<int> job_stream;  // stream index of each job
    uint64_t mbs = 0, inter_mbs = 0, intra_mbs = 0, parts = 0, coefs = 0, ops = 0, inter_coefs = 0;
};

struct Staged {
    uint8_t* d = nullptr;
    StepLayout L;
};

// One in-flight result of the pipelined path: converted on the device, copied to pinned host memory.
struct OutSlot {
    uint8_t* h = nullptr;
    uint8_t* d = nullptr;
    const uint8_t** ptr_h = nullptr;
    const uint8_t** ptr_d = nullptr;
    size_t cap = 0, bytes = 0;
    cudaEvent_t ready = nullptr, converted = nullptr;
};

}  // namespace

// Caller-supplied packed arrays are untrusted: everything the kernels index with is range-checked here (host only; also
// exported as mobi_packed_validate so that it can be exercised without a device).
struct PackedValidator {
    Geom g_;
    uint32_t H_, n_mb_;
    std::string err_;
    PackedValidator(const Geom& g, uint32_t h) : g_(g), H_(h), n_mb_((uint32_t)(g.mbw * g.mbh)) {}
    int set_err(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err_ = buf;
        return code;
    }
    int run(const mobi_packed_frame& f, int pictures) {
        const mobi_frame_hdr& h = *f.hdr;
        if (h.n_mb != n_mb_) return set_err(MOBI_ERR_ARG, "packed frame: n_mb %u, geometry needs %u", h.n_mb, n_mb_);
        if ((h.n_mb && !f.mbs) || (h.n_parts && !f.parts) || (h.n_ops && !f.ops) || (h.n_coefs && !f.coefs) || (h.n_intra && !f.intra_list))
            return set_err(MOBI_ERR_ARG, "packed frame: null array");
        const int S = g_.S, H = (int)H_;
        uint32_t n_intra = 0;
        for (uint32_t m = 0; m < h.n_mb; m++) {
            const mobi_mb& mb = f.mbs[m];
            const uint32_t kind = mb.info & 3u, nsub = (mb.info >> 2) & 127u, nco = (mb.info >> 9) & 511u;
            if (kind > 1 || nsub > 64 || nco > 384) return set_err(MOBI_ERR_ARG, "packed frame: MB %u descriptor", m);
            if ((uint64_t)mb.first_coef + nco > h.n_coefs) return set_err(MOBI_ERR_ARG, "packed frame: MB %u coefficient range", m);
            // inter macroblocks: which coded blocks are 8x8-transformed (info bits 24-27, 29-30) -- the inter kernel lays the
            // coefficient pool out by this mask, so every record must agree with it and name a block the macroblock codes
            const uint32_t bmask = (mb.info >> 18) & 63u, m8 = kind == 0 ? ((mb.info >> 24) & 15u) | ((mb.info >> 29) & 3u) << 4 : 0u;
            if (kind == 0 && (m8 & ~bmask)) return set_err(MOBI_ERR_ARG, "packed frame: MB %u 8x8-transform mask names blocks that are not coded", m);
            for (uint32_t k = 0; k < nco; k++) {
                const mobi_coef& c = f.coefs[mb.first_coef + k];
                if (((c.blk >> 3) & 3u) != (m & 3u)) return set_err(MOBI_ERR_ARG, "packed frame: MB %u coefficient %u owner tag", m, k);
                if (kind == 0 && (c.blk & 7u) < 6u) {
                    if (!((bmask >> (c.blk & 7u)) & 1u)) return set_err(MOBI_ERR_ARG, "packed frame: MB %u coefficient %u names a block the macroblock does not code", m, k);
                    if (((m8 >> (c.blk & 7u)) & 1u) != (uint32_t)(c.blk >> 7)) return set_err(MOBI_ERR_ARG, "packed frame: MB %u coefficient %u disagrees with the 8x8-transform mask", m, k);
                }
            }
            // the kernels pool the coefficient records of four consecutive macroblocks as ONE range of the array
            if (m + 1 < h.n_mb && f.mbs[m + 1].first_coef != mb.first_coef + nco)
                return set_err(MOBI_ERR_ARG, "packed frame: MB %u coefficient ranges are not contiguous", m);
            if (kind == 1) {
                if (nsub > 32) return set_err(MOBI_ERR_ARG, "packed frame: MB %u has more intra ops than any macroblock can (27)", m);
                if ((uint64_t)mb.first_sub + nsub > h.n_ops) return set_err(MOBI_ERR_ARG, "packed frame: MB %u op range", m);
                // every op field the intra kernels index shared memory with (tiles are 16x16 luma / 8x8 chroma + a halo)
                const int mbx = (int)(m % (uint32_t)g_.mbw), mby = (int)(m / (uint32_t)g_.mbw);
                const int mboff = mby * 16 * S + mbx * 16;
                const uint32_t mask = (mb.info >> 18) & 63u;
                for (uint32_t k = 0; k < nsub; k++) {
                    const uint32_t op = f.ops[mb.first_sub + k];
                    const uint32_t mode = op & 31u, plane = (op >> 6) & 3u, x4 = (op >> 8) & 3u, y4 = (op >> 10) & 3u;
                    const bool res = (op >> 5) & 1u;
                    if (mode > 20 || plane > 2) return set_err(MOBI_ERR_ARG, "packed frame: MB %u op %u mode / plane", m, k);
                    // decode order (MD:1759-1880): the luma blocks, then chroma; the I-picture kernel's luma and chroma warps each walk their own range
                    if (k && plane == 0 && ((f.ops[mb.first_sub + k - 1] >> 6) & 3u) != 0) return set_err(MOBI_ERR_ARG, "packed frame: MB %u op %u: luma ops must precede chroma ops", m, k);
                    const uint32_t n4 = mode == 20 ? 4u : mode >= 10 ? 1u : 2u, cells = plane == 0 ? 4u : 2u;   // block and plane width in 4-pixel cells
                    if (x4 + n4 > cells || y4 + n4 > cells) return set_err(MOBI_ERR_ARG, "packed frame: MB %u op %u block outside the macroblock", m, k);
                    if (plane != 0 && (mode == 8 || mode == 18)) return set_err(MOBI_ERR_ARG, "packed frame: MB %u op %u: chroma has no predictor 8 (MD:1866)", m, k);
                    if (mode == 20 && res) return set_err(MOBI_ERR_ARG, "packed frame: MB %u op %u: the 16x16 plane predictor carries no residual", m, k);
                    const uint32_t blk = plane == 0 ? (y4 >> 1) * 2u + (x4 >> 1) : 3u + plane;
                    if (res && !((mask >> blk) & 1u)) return set_err(MOBI_ERR_ARG, "packed frame: MB %u op %u names a residual its macroblock does not code", m, k);
                    // what the parser's check_intra_reads rejects: a predictor reading above / left of the plane (MD:1893 etc. would throw)
                    const int off = plane == 0 ? mboff + (int)y4 * 4 * S + (int)x4 * 4 : mboff / 2 + (plane == 2 ? S / 2 : 0) + (int)y4 * 4 * S + (int)x4 * 4;
                    int lo = 0; bool reads = true;
                    switch (mode) {
                    case 0: case 8: case 2: case 10: case 18: case 12: case 20: lo = off - S; break;
                    case 1: case 4: case 11: case 14: lo = off - 1; break;
                    case 5: case 6: case 7: case 15: case 16: case 17: lo = off - S - 1; break;
                    default: reads = false; break;
                    }
                    if (reads && (mode == 2 || mode == 12 || mode == 20) && off - 1 < lo) lo = off - 1;
                    if (reads && lo < 0) return set_err(MOBI_ERR_RANGE, "packed frame: MB %u op %u predictor reads above/left of the picture", m, k);
                }
                if (mb.intra_rank != n_intra || mb.intra_rank >= h.n_intra || f.intra_list[mb.intra_rank] != m)   // ranks ascend in raster order
                    return set_err(MOBI_ERR_ARG, "packed frame: MB %u intra rank", m);
                n_intra++;
                continue;
            }
            if (nsub == 0 || (uint64_t)mb.first_sub + nsub > h.n_parts) return set_err(MOBI_ERR_ARG, "packed frame: MB %u partition range", m);
            if ((mb.info >> 28) & 1u) {  // inline copy of the single partition must agree with the record that is range-checked below
                const mobi_part& p = f.parts[mb.first_sub];
                const uint32_t want = ((uint32_t)p.mvx & 0x3FFFu) | ((uint32_t)p.mvy & 0x3FFFu) << 14 | (uint32_t)(p.shape >> 4) << 28;
                if (nsub != 1 || mb.intra_rank != want || p.mvx < -8192 || p.mvx >= 8192 || p.mvy < -8192 || p.mvy >= 8192)
                    return set_err(MOBI_ERR_ARG, "packed frame: MB %u inline partition", m);
            }
            const int mbx = (int)(m % (uint32_t)g_.mbw), mby = (int)(m / (uint32_t)g_.mbw);
            uint64_t cover[4] = {0, 0, 0, 0};  // 16x16 luma pixels
            for (uint32_t k = 0; k < nsub; k++) {
                const mobi_part& p = f.parts[mb.first_sub + k];
                const int x = (p.xy & 15) * 2, y = (p.xy >> 4) * 2, w = 2 << (p.shape & 3), hh = 2 << ((p.shape >> 2) & 3), ref = p.shape >> 4;
                if (x + w > 16 || y + hh > 16) return set_err(MOBI_ERR_ARG, "packed frame: MB %u partition geometry", m);
                if (ref < 1 || ref > 5 || ref > pictures) return set_err(MOBI_ERR_REFERENCE, "packed frame: MB %u references picture %d of %d", m, ref, pictures);
                const long long offp = (long long)(mby * 16 + y) * S + mbx * 16 + x;
                const int dx = p.mvx, dy = p.mvy;
                const long long first = offp + (long long)(dy >> 1) * S + (dx >> 1);
                const long long last = first + (long long)(hh - 1 + (dy & 1)) * S + w - 1 + (dx & 1);
                if (first < 0 || last >= (long long)S * H) return set_err(MOBI_ERR_RANGE, "packed frame: MB %u luma vector", m);
                const int cdx = dx >> 1, cdy = dy >> 1;
                const long long cfirst = offp / 2 + (long long)(cdy >> 1) * S + (cdx >> 1);
                const long long clast = cfirst + S / 2 + (long long)((hh >> 1) - 1 + (cdy & 1)) * S + (w >> 1) - 1 + (cdx & 1);
                if (cfirst < 0 || clast >= (long long)S * H / 2) return set_err(MOBI_ERR_RANGE, "packed frame: MB %u chroma vector", m);
                for (int yy = y; yy < y + hh; yy++) cover[yy >> 2] |= (uint64_t)((1u << w) - 1u) << ((yy & 3) * 16 + x);
            }
            for (int k = 0; k < 4; k++) if (cover[k] != ~0ull) return set_err(MOBI_ERR_ARG, "packed frame: MB %u partitions do not tile the macroblock", m);
        }
        if (n_intra != h.n_intra) return set_err(MOBI_ERR_ARG, "packed frame: intra count");
        for (uint32_t k = 0; k < h.n_coefs; k++) {
            const mobi_coef& c = f.coefs[k];
            if ((c.blk & 7) > 5) return set_err(MOBI_ERR_ARG, "packed frame: coefficient %u block tag", k);
            if (!(c.blk & 0x80) && (c.pos & 63) > 15) return set_err(MOBI_ERR_ARG, "packed frame: coefficient %u scan position", k);
        }
        for (int k = 0; k < 80; k++) {
            const uint32_t idx = h.qtab[k] & 0xFF;
            if (idx >= (k < 64 ? 64u : 16u)) return set_err(MOBI_ERR_ARG, "packed frame: scan table entry %d", k);
        }
        return MOBI_OK;
    }

};

class Batch {
public:
    Batch(uint32_t w, uint32_t h, int version, int device, int n_streams, int n_threads)
        : W_(w), H_(h), ver_(version), dev_(device), N_(n_streams), pool_(n_threads > 1 ? n_threads - 1 : 0) {
        g_.W = (int)w; g_.H = (int)h; g_.S = stride_for(w); g_.version = version;
        g_.log2S = g_.S == 256 ? 8 : g_.S == 512 ? 9 : 10;
        g_.mbw = (int)w / 16; g_.mbh = (int)h / 16;
        n_mb_ = (uint32_t)(g_.mbw * g_.mbh);
        ysz_ = (size_t)g_.S * h;
        pic_ = align_up(ysz_ * 3 / 2 + 256, (size_t)g_.S);   // whole rows: the ring is also addressed as one tensor of rows
        for (int i = 0; i < n_streams; i++) parsers_.emplace_back(new Parser(w, h, version));
        frames_.resize(n_streams);
        count_.assign(n_streams, 0);
        staged_count_.assign(n_streams, 0);
    }
    ~Batch() {
        if (getenv("MOBI_PACK_TRACE")) fprintf(stderr, "[mobicuda] pack: arena wait %.1f ms, copies %.1f ms, heights %.1f ms, sort %.1f ms (totals since creation; device %d)\n", pack_ms_[0], pack_ms_[1], pack_ms_[2], pack_ms_[3], dev_);
        if (stream_) {
            cudaSetDevice(dev_);
            cudaStreamSynchronize(stream_);
            clear_staged();
            for (auto& a : arena_) { if (a.h) cudaFreeHost(a.h); if (a.d) cudaFree(a.d); if (a.consumed) cudaEventDestroy(a.consumed); }
            if (ring_) cudaFree(ring_);
            if (flags_) cudaFree(flags_);
            if (ticket_) cudaFree(ticket_);
            if (out_d_) cudaFree(out_d_);
            if (out_h_) cudaFreeHost(out_h_);
            if (ptr_d_) cudaFree(ptr_d_);
            if (ring_tab_d_) cudaFree(ring_tab_d_);
            if (ptr_h_) cudaFreeHost(ptr_h_);
            for (int w = 0; w < 5; w++) for (cudaEvent_t e : ev_[w]) cudaEventDestroy(e);
            if (side_) { cudaStreamSynchronize(side_); cudaStreamDestroy(side_); }
            if (copy_) { cudaStreamSynchronize(copy_); cudaStreamDestroy(copy_); }
            if (conv_) { cudaStreamSynchronize(conv_); cudaStreamDestroy(conv_); }
            if (conv_go_) cudaEventDestroy(conv_go_);
            if (conv_done_) cudaEventDestroy(conv_done_);
            if (fork_) cudaEventDestroy(fork_);
            if (join_) cudaEventDestroy(join_);
            for (cudaEvent_t e : ev_free_) cudaEventDestroy(e);
            for (auto& o : slot_) {
                if (o.h) cudaFreeHost(o.h);
                if (o.d) cudaFree(o.d);
                if (o.ptr_h) cudaFreeHost(o.ptr_h);
                if (o.ptr_d) cudaFree(o.ptr_d);
                if (o.ready) cudaEventDestroy(o.ready);
                if (o.converted) cudaEventDestroy(o.converted);
            }
            cudaStreamDestroy(stream_);
        }
    }

    int init() {
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        cudaDeviceProp prop;
        if (!ok(cudaGetDeviceProperties(&prop, dev_), "cudaGetDeviceProperties")) return MOBI_ERR_CUDA;
        sm_count_ = prop.multiProcessorCount;
        key_ticket_threshold_ = sm_count_;
        if (const char* e = getenv("MOBI_KEY_TICKETS")) key_ticket_threshold_ = atoi(e);   // experiments: 0 = always by ticket, a huge value = never
        if (!ok(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate")) return MOBI_ERR_CUDA;
        if (!ok(cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking), "cudaStreamCreate")) return MOBI_ERR_CUDA;
        if (!ok(cudaStreamCreateWithFlags(&copy_, cudaStreamNonBlocking), "cudaStreamCreate")) return MOBI_ERR_CUDA;
        if (!ok(cudaEventCreateWithFlags(&fork_, cudaEventDisableTiming), "cudaEventCreate")) return MOBI_ERR_CUDA;
        if (!ok(cudaEventCreateWithFlags(&join_, cudaEventDisableTiming), "cudaEventCreate")) return MOBI_ERR_CUDA;
        if (!ok(cudaMalloc(&ring_, pic_ * RING * (size_t)N_), "cudaMalloc(ring)")) return MOBI_ERR_NOMEM;
        if (!ok(cudaMemsetAsync(ring_, 0, pic_ * RING * (size_t)N_, stream_), "memset(ring)")) return MOBI_ERR_CUDA;
        if (!ok(cudaMalloc(&flags_, sizeof(uint32_t) * n_mb_ * (size_t)N_), "cudaMalloc(flags)")) return MOBI_ERR_NOMEM;
        if (!ok(cudaMemsetAsync(flags_, 0, sizeof(uint32_t) * n_mb_ * (size_t)N_, stream_), "memset(flags)")) return MOBI_ERR_CUDA;
        if (!ok(cudaMalloc(&ticket_, 256), "cudaMalloc(ticket)")) return MOBI_ERR_NOMEM;
        if (!ok(cudaMemsetAsync(ticket_, 0, 256, stream_), "memset(ticket)")) return MOBI_ERR_CUDA;
        for (auto& a : arena_)
            if (!ok(cudaEventCreateWithFlags(&a.consumed, cudaEventDisableTiming), "cudaEventCreate")) return MOBI_ERR_CUDA;
        if (!ok(cudaMalloc(&ptr_d_, sizeof(void*) * (size_t)N_), "cudaMalloc(ptrs)")) return MOBI_ERR_NOMEM;
        if (!ok(cudaMallocHost(&ptr_h_, sizeof(void*) * (size_t)N_), "cudaMallocHost(ptrs)")) return MOBI_ERR_NOMEM;
        {   // every ring position's address, resident: a single picture is converted without uploading its pointer first
            std::vector<const uint8_t*> tab((size_t)N_ * RING);
            for (int s = 0; s < N_; s++) for (int k = 0; k < RING; k++) tab[(size_t)s * RING + k] = picture(s, k);
            if (!ok(cudaMalloc(&ring_tab_d_, sizeof(void*) * tab.size()), "cudaMalloc(ring table)")) return MOBI_ERR_NOMEM;
            if (!ok(cudaMemcpy(ring_tab_d_, tab.data(), sizeof(void*) * tab.size(), cudaMemcpyHostToDevice), "H2D ring table")) return MOBI_ERR_CUDA;
        }
        if (!ok(init_kernel_tables(), "init_kernel_tables")) return MOBI_ERR_CUDA;
        if (!ok(cudaStreamSynchronize(stream_), "init sync")) return MOBI_ERR_CUDA;
        return make_tensor_maps();
    }
    // The ring as TMA tensors (InterMaps, mobi_kernels.h).  cuTensorMapEncodeTiled is reached through the runtime so that the
    // library does not link libcuda.
    int make_tensor_maps() {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (!ok(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres), "cudaGetDriverEntryPoint") || !fn || qres != cudaDriverEntryPointSuccess)
            return set_err(MOBI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
        const cuuint64_t S = (cuuint64_t)g_.S, rows = (cuuint64_t)(pic_ / S), pics = (cuuint64_t)N_ * RING;
        tm_.ring_rows = (int)rows;
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        // MOBI_TMA_L2PROMO = 64 / 128 / 256: L2 promotion of the boxes' misses (experiments; default none)
        CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;
        if (const char* e = getenv("MOBI_TMA_L2PROMO")) {
            const int v = atoi(e);
            promo = v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo;
        }
        auto enc = [&](CUtensorMap* m, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* bx) {
            return encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, ring_, dims, strides, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        };
        CUresult r;
        {
            const cuuint64_t dims[2] = {S, rows * pics}, strides[1] = {S};
            const cuuint32_t bx[2] = {32, 17};
            r = enc(&tm_.l2, 2, dims, strides, bx);
        }
        if (r == CUDA_SUCCESS) {
            const cuuint64_t dims[3] = {S / 2, 2, rows * pics}, strides[2] = {S / 2, S};
            const cuuint32_t bx[3] = {32, 2, 9};
            r = enc(&tm_.c3, 3, dims, strides, bx);
        }
        if (r == CUDA_SUCCESS) {
            const cuuint64_t dims[3] = {S, (cuuint64_t)H_ * 3 / 2, pics}, strides[2] = {S, (cuuint64_t)pic_};
            const cuuint32_t bx[3] = {32, 17, 1};
            r = enc(&tm_.l3, 3, dims, strides, bx);
        }
        if (r == CUDA_SUCCESS) {
            const cuuint64_t dims[4] = {S / 2, 2, (cuuint64_t)H_ * 3 / 2, pics}, strides[3] = {S / 2, S, (cuuint64_t)pic_};
            const cuuint32_t bx[4] = {32, 2, 9, 1};
            r = enc(&tm_.c4, 4, dims, strides, bx);
        }
        if (r != CUDA_SUCCESS) return set_err(MOBI_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        return MOBI_OK;
    }

    // ---- one lock-step advance from raw frame bytes ---------------------------------------------
    int decode(const uint8_t* const* data, const int* len, int* offset, int* status) {
        if (!data || !len || !offset) return set_err(MOBI_ERR_ARG, "null argument");
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        rc_.assign(N_, MOBI_OK);
        const auto t0 = std::chrono::steady_clock::now();
        pool_.run(N_, [&](int i) { rc_[i] = parse_one(i, data[i], len[i], &offset[i], count_[i]); });
        const auto t1 = std::chrono::steady_clock::now();
        phase_ms_[0] += std::chrono::duration<double, std::milli>(t1 - t0).count();
        views_.resize(N_);
        int n_ok = 0, first_bad = -1;
        for (int i = 0; i < N_; i++) {
            if (status) status[i] = rc_[i];
            if (rc_[i] == MOBI_OK) { views_[i] = frames_[i].view(); n_ok++; }
            else { views_[i] = mobi_packed_frame{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; if (first_bad < 0) first_bad = i; }
        }
        if (first_bad >= 0) set_err(rc_[first_bad], "stream %d: %s", first_bad, parsers_[first_bad]->error().c_str());
        if (n_ok == 0) return first_bad >= 0 ? rc_[first_bad] : MOBI_OK;
        Arena& a = arena_[cur_arena_];
        cur_arena_ ^= 1;
        StepLayout L;
        int rc = pack_step(a, L, count_);
        const auto t2 = std::chrono::steady_clock::now();
        phase_ms_[1] += std::chrono::duration<double, std::milli>(t2 - t1).count();
        if (rc != MOBI_OK) return rc;
        if (!ok(cudaMemcpyAsync(a.d, a.h, L.bytes, cudaMemcpyHostToDevice, stream_), "H2D arena")) return MOBI_ERR_CUDA;
        stats_.h2d_bytes += L.bytes;
        rc = launch_step(a.d, L);
        if (rc != MOBI_OK) return rc;
        cudaEventRecord(a.consumed, stream_);
        a.pending = true;
        for (int s : L.job_stream) count_[s]++;
        phase_ms_[2] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t2).count();
        return (first_bad >= 0 && N_ == 1) ? rc_[first_bad] : MOBI_OK;
    }

    // Reconstruct stream 0's next picture from caller-provided packed arrays (mobi_submit_packed).
    int submit(int stream, const mobi_packed_frame* f) {
        if (!f || !f->hdr) return set_err(MOBI_ERR_ARG, "null packed frame");
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        int rc = validate(*f, count_[stream]);
        if (rc != MOBI_OK) return rc;
        views_.assign(N_, mobi_packed_frame{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr});
        views_[stream] = *f;
        Arena& a = arena_[cur_arena_];
        cur_arena_ ^= 1;
        StepLayout L;
        rc = pack_step(a, L, count_);
        if (rc != MOBI_OK) return rc;
        if (!ok(cudaMemcpyAsync(a.d, a.h, L.bytes, cudaMemcpyHostToDevice, stream_), "H2D arena")) return MOBI_ERR_CUDA;
        stats_.h2d_bytes += L.bytes;
        rc = launch_step(a.d, L);
        if (rc != MOBI_OK) return rc;
        cudaEventRecord(a.consumed, stream_);
        a.pending = true;
        count_[stream]++;
        quant_override_ = f->hdr->quantizer; yuv_override_ = f->hdr->yuv_format; have_override_ = true;
        return MOBI_OK;
    }

    // ---- staging / replay -------------------------------------------------------------------------
    int stage(const uint8_t* const* data, const int* len, int* offset) {
        if (!data || !len || !offset) return set_err(MOBI_ERR_ARG, "null argument");
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        rc_.assign(N_, MOBI_OK);
        pool_.run(N_, [&](int i) { rc_[i] = parse_one(i, data[i], len[i], &offset[i], staged_count_[i]); });
        views_.resize(N_);
        for (int i = 0; i < N_; i++) {
            if (rc_[i] != MOBI_OK) return set_err(rc_[i], "stream %d: %s", i, parsers_[i]->error().c_str());
            views_[i] = frames_[i].view();
        }
        Arena& a = arena_[cur_arena_];
        cur_arena_ ^= 1;
        Staged st;
        int rc = pack_step(a, st.L, staged_count_);
        if (rc != MOBI_OK) return rc;
        if (!ok(cudaMalloc(&st.d, st.L.bytes), "cudaMalloc(staged step)")) return MOBI_ERR_NOMEM;
        // the job table holds device addresses relative to the arena's device buffer: rebase onto st.d
        rebase(a.h, st.L, a.d, st.d);
        if (!ok(cudaMemcpyAsync(st.d, a.h, st.L.bytes, cudaMemcpyHostToDevice, stream_), "H2D staged")) return MOBI_ERR_CUDA;
        cudaEventRecord(a.consumed, stream_);
        a.pending = true;
        for (int s : st.L.job_stream) staged_count_[s]++;
        staged_.push_back(std::move(st));
        return MOBI_OK;
    }
    // format 0: reconstruction only; MOBI_OUT_BGRA: every step's new pictures are also converted (k_bgra) into the device-side
    // output buffer.  Nothing crosses PCIe either way.
    int replay(int first, int count, int format = 0) {
        if (first < 0 || count < 0 || first + count > (int)staged_.size()) return set_err(MOBI_ERR_ARG, "replay range outside staged steps");
        if (format != 0 && format != 2) return set_err(MOBI_ERR_ARG, "replay converts to BGRA (2) or not at all (0)");
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        const size_t per = (size_t)W_ * H_ * 4;
        if (format) {
            int rc = ensure_out(per * N_);
            if (rc != MOBI_OK) return rc;
            if (!conv_ && !ok(cudaStreamCreateWithFlags(&conv_, cudaStreamNonBlocking), "cudaStreamCreate")) return MOBI_ERR_CUDA;
            if (!conv_go_ && (!ok(cudaEventCreateWithFlags(&conv_go_, cudaEventDisableTiming), "cudaEventCreate") ||
                              !ok(cudaEventCreateWithFlags(&conv_done_, cudaEventDisableTiming), "cudaEventCreate"))) return MOBI_ERR_CUDA;
        }
        for (int k = first; k < first + count; k++) {
            int rc = launch_step(staged_[k].d, staged_[k].L);
            if (rc != MOBI_OK) return rc;
            for (int s : staged_[k].L.job_stream) count_[s]++;
            if (format) {
                const StepLayout& L = staged_[k].L;
                // MOBI_BGRA_STREAM=own puts the conversion on a stream of its own, beside the next step's reconstruction.  Measured
                // (B200, 1024 x 400x240): the two then fight for the same SMs and memory system -- k_bgra 0.38 ms instead of 0.12,
                // the inter kernel 0.60 instead of 0.26, 0.70 ms per step instead of 0.48 -- so the default is back to back.
                static const bool own = [] { const char* e = getenv("MOBI_BGRA_STREAM"); return e && !strcmp(e, "own"); }();
                cudaStream_t cs = own ? conv_ : stream_;
                if (own) {
                    if (!ok(cudaEventRecord(conv_go_, stream_), "cudaEventRecord")) return MOBI_ERR_CUDA;
                    if (!ok(cudaStreamWaitEvent(conv_, conv_go_, 0), "cudaStreamWaitEvent")) return MOBI_ERR_CUDA;
                }
                if (timing_) tick(4, cs);
                if (!ok(launch_bgra(reinterpret_cast<const uint8_t* const*>(staged_[k].d + L.pics_off), L.n_jobs, out_d_, (int)W_ * 4, per, g_, cs), "k_bgra")) return MOBI_ERR_CUDA;
                if (timing_) tick(4, cs);
                stats_.launches++;
            }
        }
        if (format && count > 0) {   // whoever waits for the batch's stream waits for the last conversion too
            if (!ok(cudaEventRecord(conv_done_, conv_), "cudaEventRecord")) return MOBI_ERR_CUDA;
            if (!ok(cudaStreamWaitEvent(stream_, conv_done_, 0), "cudaStreamWaitEvent")) return MOBI_ERR_CUDA;
        }
        return MOBI_OK;
    }
    int staged_steps() const { return (int)staged_.size(); }
    void clear_staged() {
        if (stream_) cudaStreamSynchronize(stream_);
        for (auto& s : staged_) cudaFree(s.d);
        staged_.clear();
        std::fill(staged_count_.begin(), staged_count_.end(), 0);
    }
    // Ring back to "nothing decoded".  Staged steps stay valid: they were laid out from picture 0.
    int reset(bool parsers_too) {
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        if (!ok(cudaStreamSynchronize(stream_), "sync")) return MOBI_ERR_CUDA;
        std::fill(count_.begin(), count_.end(), 0);
        if (parsers_too) { for (auto& p : parsers_) p->reset(); std::fill(staged_count_.begin(), staged_count_.end(), 0); }
        have_override_ = false;
        return MOBI_OK;
    }
    int sync() {
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        if (!ok(cudaStreamSynchronize(stream_), "cudaStreamSynchronize")) return MOBI_ERR_CUDA;
        return ok(cudaStreamSynchronize(copy_), "cudaStreamSynchronize") ? MOBI_OK : MOBI_ERR_CUDA;
    }

    // ---- read-back --------------------------------------------------------------------------------
    int read_strided(int s, uint8_t* y, uint8_t* uv) {
        if (s < 0 || s >= N_) return set_err(MOBI_ERR_ARG, "stream index");
        if (count_[s] == 0) return set_err(MOBI_ERR_STATE, "no picture decoded yet");
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        const uint8_t* p = picture(s, count_[s] - 1);
        if (y && !ok(cudaMemcpyAsync(y, p, ysz_, cudaMemcpyDeviceToHost, stream_), "D2H luma")) return MOBI_ERR_CUDA;
        if (uv && !ok(cudaMemcpyAsync(uv, p + ysz_, ysz_ / 2, cudaMemcpyDeviceToHost, stream_), "D2H chroma")) return MOBI_ERR_CUDA;
        stats_.d2h_bytes += (y ? ysz_ : 0) + (uv ? ysz_ / 2 : 0);
        return sync();
    }
    // tight I420 of every stream's newest picture; dst may be null (device-side pack only)
    int read_yuv_all(uint8_t* dst) {
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        const size_t per = (size_t)W_ * H_ * 3 / 2, total = per * N_;
        int rc = ensure_out(total);
        if (rc != MOBI_OK) return rc;
        for (int s = 0; s < N_; s++) {
            if (count_[s] == 0) return set_err(MOBI_ERR_STATE, "stream %d: no picture decoded yet", s);
            ptr_h_[s] = picture(s, count_[s] - 1);
        }
        // ptr_h_ is re-used every call: the previous call's copy has completed because every read syncs
        if (!ok(cudaMemcpyAsync(ptr_d_, ptr_h_, sizeof(void*) * N_, cudaMemcpyHostToDevice, stream_), "H2D ptrs")) return MOBI_ERR_CUDA;
        if (!ok(launch_pack_i420(ptr_d_, N_, out_d_, g_, stream_), "k_pack_i420")) return MOBI_ERR_CUDA;
        stats_.launches++;
        if (dst) {
            if (!ok(cudaMemcpyAsync(out_h_, out_d_, total, cudaMemcpyDeviceToHost, stream_), "D2H i420")) return MOBI_ERR_CUDA;
            stats_.d2h_bytes += total;
        }
        rc = sync();
        if (rc != MOBI_OK) return rc;
        if (dst) {
            // pinned staging -> caller memory (pageable in general), fanned out over the pool
            const int chunks = N_;
            pool_.run(chunks, [&](int i) { std::memcpy(dst + per * i, out_h_ + per * i, per); });
        }
        return MOBI_OK;
    }
    int read_yuv_one(int s, uint8_t* y, uint8_t* u, uint8_t* v) {
        if (s < 0 || s >= N_) return set_err(MOBI_ERR_ARG, "stream index");
        if (count_[s] == 0) return set_err(MOBI_ERR_STATE, "no picture decoded yet");
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        const uint8_t* p = picture(s, count_[s] - 1);
        const size_t S = (size_t)g_.S;
        if (y && !ok(cudaMemcpy2DAsync(y, W_, p, S, W_, H_, cudaMemcpyDeviceToHost, stream_), "D2H Y")) return MOBI_ERR_CUDA;
        if (u && !ok(cudaMemcpy2DAsync(u, W_ / 2, p + ysz_, S, W_ / 2, H_ / 2, cudaMemcpyDeviceToHost, stream_), "D2H U")) return MOBI_ERR_CUDA;
        if (v && !ok(cudaMemcpy2DAsync(v, W_ / 2, p + ysz_ + S / 2, S, W_ / 2, H_ / 2, cudaMemcpyDeviceToHost, stream_), "D2H V")) return MOBI_ERR_CUDA;
        stats_.d2h_bytes += (size_t)W_ * H_ * 3 / 2;
        return sync();
    }
    // BGRA of every stream's newest picture into one device buffer; optionally down to the host.
    int bgra_all(uint8_t* dst) {
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        const size_t per = (size_t)W_ * H_ * 4, total = per * N_;
        int rc = ensure_out(total);
        if (rc != MOBI_OK) return rc;
        for (int s = 0; s < N_; s++) {
            if (count_[s] == 0) return set_err(MOBI_ERR_STATE, "stream %d: no picture decoded yet", s);
            ptr_h_[s] = picture(s, count_[s] - 1);
        }
        if (!ok(cudaMemcpyAsync(ptr_d_, ptr_h_, sizeof(void*) * N_, cudaMemcpyHostToDevice, stream_), "H2D ptrs")) return MOBI_ERR_CUDA;
        if (!ok(launch_bgra(ptr_d_, N_, out_d_, (int)W_ * 4, per, g_, stream_), "k_bgra")) return MOBI_ERR_CUDA;
        stats_.launches++;
        if (dst) {
            if (!ok(cudaMemcpyAsync(out_h_, out_d_, total, cudaMemcpyDeviceToHost, stream_), "D2H bgra")) return MOBI_ERR_CUDA;
            stats_.d2h_bytes += total;
        }
        rc = sync();
        if (rc != MOBI_OK) return rc;
        if (dst) pool_.run(N_, [&](int i) { std::memcpy(dst + per * i, out_h_ + per * i, per); });
        return MOBI_OK;
    }
    int read_bgra_one(int s, uint8_t* dst, int dst_stride) {
        if (s < 0 || s >= N_ || !dst || dst_stride < (int)W_ * 4) return set_err(MOBI_ERR_ARG, "bad bgra destination");
        if (count_[s] == 0) return set_err(MOBI_ERR_STATE, "no picture decoded yet");
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        const size_t per = (size_t)W_ * H_ * 4;
        int rc = ensure_out(per);
        if (rc != MOBI_OK) return rc;
        if (!ok(launch_bgra(ring_tab_d_ + (size_t)s * RING + (size_t)((count_[s] - 1) % RING), 1, out_d_, (int)W_ * 4, per, g_, stream_), "k_bgra")) return MOBI_ERR_CUDA;
        stats_.launches++;
        if (!ok(cudaMemcpy2DAsync(dst, (size_t)dst_stride, out_d_, (size_t)W_ * 4, (size_t)W_ * 4, H_, cudaMemcpyDeviceToHost, stream_), "D2H bgra")) return MOBI_ERR_CUDA;
        stats_.d2h_bytes += per;
        return sync();
    }

    // ---- pipelined path: submit step k+1 while step k is still on the GPU ---------------------------
    // format: 1 = tight I420, 2 = BGRA.  At most two results may be outstanding.
    int submit_async(const uint8_t* const* data, const int* len, int* offset, int* status, int format) {
        if (format != 1 && format != 2) return set_err(MOBI_ERR_ARG, "output format must be 1 (I420) or 2 (BGRA)");
        if (slots_used_ == 2) return set_err(MOBI_ERR_STATE, "two results outstanding: fetch one first");
        int rc = decode(data, len, offset, status);
        if (rc != MOBI_OK) return rc;
        OutSlot& o = slot_[slot_head_];
        const size_t per = format == 1 ? (size_t)W_ * H_ * 3 / 2 : (size_t)W_ * H_ * 4, total = per * N_;
        if (o.cap < total) {
            if (o.h) cudaFreeHost(o.h);
            if (o.d) cudaFree(o.d);
            o.h = nullptr; o.d = nullptr; o.cap = 0;
            if (!ok(cudaMallocHost(&o.h, total), "cudaMallocHost(result)")) return MOBI_ERR_NOMEM;
            if (!ok(cudaMalloc(&o.d, total), "cudaMalloc(result)")) return MOBI_ERR_NOMEM;
            o.cap = total;
        }
        if (!o.ptr_h) {
            if (!ok(cudaMallocHost(&o.ptr_h, sizeof(void*) * N_), "cudaMallocHost(ptrs)")) return MOBI_ERR_NOMEM;
            if (!ok(cudaMalloc(&o.ptr_d, sizeof(void*) * N_), "cudaMalloc(ptrs)")) return MOBI_ERR_NOMEM;
            if (!ok(cudaEventCreateWithFlags(&o.ready, cudaEventDisableTiming), "cudaEventCreate")) return MOBI_ERR_CUDA;
            if (!ok(cudaEventCreateWithFlags(&o.converted, cudaEventDisableTiming), "cudaEventCreate")) return MOBI_ERR_CUDA;
        }
        for (int s = 0; s < N_; s++) {
            if (count_[s] == 0) return set_err(MOBI_ERR_STATE, "stream %d: no picture decoded yet", s);
            o.ptr_h[s] = picture(s, count_[s] - 1);
        }
        if (!ok(cudaMemcpyAsync(o.ptr_d, o.ptr_h, sizeof(void*) * N_, cudaMemcpyHostToDevice, stream_), "H2D ptrs")) return MOBI_ERR_CUDA;
        if (format == 1) { if (!ok(launch_pack_i420(o.ptr_d, N_, o.d, g_, stream_), "k_pack_i420")) return MOBI_ERR_CUDA; }
        else { if (!ok(launch_bgra(o.ptr_d, N_, o.d, (int)W_ * 4, per, g_, stream_), "k_bgra")) return MOBI_ERR_CUDA; }
        stats_.launches++;
        // the copy back runs on its own stream so that it overlaps the upload and the kernels of the next step
        if (!ok(cudaEventRecord(o.converted, stream_), "cudaEventRecord")) return MOBI_ERR_CUDA;
        if (!ok(cudaStreamWaitEvent(copy_, o.converted, 0), "cudaStreamWaitEvent")) return MOBI_ERR_CUDA;
        if (!ok(cudaMemcpyAsync(o.h, o.d, total, cudaMemcpyDeviceToHost, copy_), "D2H result")) return MOBI_ERR_CUDA;
        stats_.d2h_bytes += total;
        stats_.h2d_bytes += sizeof(void*) * N_;
        if (!ok(cudaEventRecord(o.ready, copy_), "cudaEventRecord")) return MOBI_ERR_CUDA;
        o.bytes = total;
        slot_head_ ^= 1;
        slots_used_++;
        return MOBI_OK;
    }
    // Oldest outstanding result.  dst != null: copied there; *view (optional) receives the pinned buffer, valid
    // until the second submit_async from now.
    int fetch(uint8_t* dst, const uint8_t** view, size_t* bytes) {
        if (slots_used_ == 0) return set_err(MOBI_ERR_STATE, "no result outstanding");
        if (!ok(cudaSetDevice(dev_), "cudaSetDevice")) return MOBI_ERR_CUDA;
        OutSlot& o = slot_[(slot_head_ + 2 - slots_used_) & 1];
        const auto t0 = std::chrono::steady_clock::now();
        if (!ok(cudaEventSynchronize(o.ready), "cudaEventSynchronize")) return MOBI_ERR_CUDA;
        phase_ms_[3] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        slots_used_--;
        if (dst) {
            const size_t per = o.bytes / N_;
            pool_.run(N_, [&](int i) { std::memcpy(dst + per * i, o.h + per * i, per); });
        }
        if (view) *view = o.h;
        if (bytes) *bytes = o.bytes;
        return MOBI_OK;
    }

    // ---- per-kernel timing (roofline accounting) ---------------------------------------------------
    void set_timing(bool on) { timing_ = on; }
    cudaEvent_t fresh_event() {
        cudaEvent_t e;
        if (ev_free_.empty()) cudaEventCreate(&e); else { e = ev_free_.back(); ev_free_.pop_back(); }
        return e;
    }
    void tick(int which, cudaStream_t st) {
        cudaEvent_t e = fresh_event();
        cudaEventRecord(e, st);
        ev_[which].push_back(e);
    }
    // which: 0 the inter kernel (k_mc; the fused kernel where one is selected), 1 k_intra over P-pictures, 2 k_intra over
    // I-pictures (side stream), 3 k_res, 4 k_bgra (replay with conversion; its own stream)
    int kernel_times(double* ms_out, uint64_t* n_out) {
        int rc = sync();
        if (rc != MOBI_OK) return rc;
        if (conv_ && !ok(cudaStreamSynchronize(conv_), "cudaStreamSynchronize")) return MOBI_ERR_CUDA;
        for (int w = 0; w < 5; w++) {
            double ms = 0; uint64_t n = 0;
            for (size_t i = 0; i + 1 < ev_[w].size(); i += 2) {
                float t = 0;
                cudaEventElapsedTime(&t, ev_[w][i], ev_[w][i + 1]);
                ms += t; n++;
            }
            for (cudaEvent_t e : ev_[w]) ev_free_.push_back(e);
            ev_[w].clear();
            if (ms_out) ms_out[w] = ms;
            if (n_out) n_out[w] = n;
        }
        return MOBI_OK;
    }

    void get_state(int s, uint32_t* q, uint32_t* yf, int* stride) const {
        if (q) *q = have_override_ ? quant_override_ : parsers_[s]->quantizer();
        if (yf) *yf = have_override_ ? yuv_override_ : parsers_[s]->yuv_format();
        if (stride) *stride = g_.S;
    }
    const char* error() const { return err_.c_str(); }
    void* cuda_stream() const { return (void*)stream_; }
    const mobi_batch_stats& stats() const { return stats_; }
    void clear_stats() { std::memset(&stats_, 0, sizeof stats_); std::memset(phase_ms_, 0, sizeof phase_ms_); }
    // host wall time since clear_stats: entropy parse (all threads, wall), packing into the pinned arena, enqueueing the
    // upload and the kernels, waiting in fetch for a result's copy-back
    void phase_times(double* ms) const { for (int i = 0; i < 4; i++) ms[i] = phase_ms_[i]; }
    int n_streams() const { return N_; }
    uint32_t width() const { return W_; }
    uint32_t height() const { return H_; }
    int set_err(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err_ = buf;
        return code;
    }

private:
    // One stream's entropy parse.  The parser is told how many pictures the ring holds for this stream first (the ring is
    // the authority: after mobi_batch_reset, or a step that failed after its parse succeeded, the parser's own count would
    // be ahead), so a P-picture naming a picture the ring does not hold is MOBI_ERR_REFERENCE here, never a null reference
    // on the device.  Nothing is thrown past this point on a worker thread.
    int parse_one(int i, const uint8_t* data, int len, int* offset, int pictures) {
        try {
            parsers_[i]->set_pictures(pictures);
            return parsers_[i]->parse(data, len, offset, frames_[i]);
        } catch (...) { return MOBI_ERR_NOMEM; }
    }
    bool ok(cudaError_t e, const char* what) {
        if (e == cudaSuccess) return true;
        set_err(MOBI_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
        return false;
    }
    uint8_t* picture(int s, int idx) const { return ring_ + ((size_t)s * RING + (size_t)(idx % RING)) * pic_; }

    int ensure_out(size_t bytes) {
        if (bytes <= out_cap_) return MOBI_OK;
        if (!ok(cudaStreamSynchronize(stream_), "sync")) return MOBI_ERR_CUDA;
        if (out_d_) cudaFree(out_d_);
        if (out_h_) cudaFreeHost(out_h_);
        out_d_ = nullptr; out_h_ = nullptr; out_cap_ = 0;
        if (!ok(cudaMalloc(&out_d_, bytes), "cudaMalloc(out)")) return MOBI_ERR_NOMEM;
        if (!ok(cudaMallocHost(&out_h_, bytes), "cudaMallocHost(out)")) return MOBI_ERR_NOMEM;
        out_cap_ = bytes;
        return MOBI_OK;
    }
    int ensure_arena(Arena& a, size_t bytes) {
        if (a.pending) {  // the GPU may still be reading the previous contents
            if (!ok(cudaEventSynchronize(a.consumed), "cudaEventSynchronize")) return MOBI_ERR_CUDA;
            a.pending = false;
        }
        if (bytes <= a.cap) return MOBI_OK;
        size_t cap = align_up(bytes + bytes / 2, 1 << 16);
        if (a.h) cudaFreeHost(a.h);
        if (a.d) cudaFree(a.d);
        a.h = nullptr; a.d = nullptr; a.cap = 0;
        if (!ok(cudaMallocHost(&a.h, cap), "cudaMallocHost(arena)")) return MOBI_ERR_NOMEM;
        if (!ok(cudaMalloc(&a.d, cap), "cudaMalloc(arena)")) return MOBI_ERR_NOMEM;
        a.cap = cap;
        return MOBI_OK;
    }

    // Lay the frames in views_ (hdr == null: stream sits this step out) into arena `a` and build the job
    // table against ring positions given by `count` (pictures decoded so far, per stream).
    int pack_step(Arena& a, StepLayout& L, const std::vector<int>& count) {
        struct Off { size_t hdr, mbs, parts, ops, coefs, intra; };
        std::vector<Off> off(N_);
        L = StepLayout();
        for (int i = 0; i < N_; i++) if (views_[i].hdr) { L.job_stream.push_back(i); }
        L.n_jobs = (int)L.job_stream.size();
        size_t p = 0;
        L.jobs_off = p; p = align_up(p + sizeof(DevJob) * L.n_jobs, 256);
        for (int j = 0; j < L.n_jobs; j++) {
            const mobi_frame_hdr& h = *views_[L.job_stream[j]].hdr;
            L.n_work += h.n_intra;
            if (h.n_intra < h.n_mb) L.n_inter_jobs++;
            L.mbs += h.n_mb; L.intra_mbs += h.n_intra; L.inter_mbs += h.n_mb - h.n_intra;
            L.parts += h.n_parts; L.coefs += h.n_coefs; L.ops += h.n_ops; L.inter_coefs += h.n_inter_coefs;
        }
        L.work_off = p; p = align_up(p + sizeof(IntraWork) * L.n_work, 256);
        L.key_off = p; p = align_up(p + 8 * (size_t)L.n_jobs, 256);
        L.pics_off = p; p = align_up(p + sizeof(void*) * (size_t)L.n_jobs, 256);
        for (int j = 0; j < L.n_jobs; j++) {
            const mobi_frame_hdr& h = *views_[L.job_stream[j]].hdr;
            Off& o = off[j];
            o.hdr = p; p = align_up(p + sizeof(mobi_frame_hdr), 16);
            o.mbs = p; p = align_up(p + sizeof(mobi_mb) * h.n_mb, 16);
            o.parts = p; p = align_up(p + sizeof(mobi_part) * h.n_parts, 16);
            o.ops = p; p = align_up(p + sizeof(mobi_op) * h.n_ops, 16);
            o.coefs = p; p = align_up(p + sizeof(mobi_coef) * h.n_coefs + 128, 16);  // +128: the intra kernel prefetches one warp past the end
            o.intra = p; p = align_up(p + sizeof(uint32_t) * h.n_intra, 256);
        }
        L.bytes = p;
        for (int j = 0; j < L.n_jobs; j++) {   // a reference the ring does not hold must never reach the device (null plane / wrong picture)
            const int s = L.job_stream[j];
            if ((int)views_[s].hdr->max_ref > count[s])
                return set_err(MOBI_ERR_REFERENCE, "stream %d: picture references ring index %u but only %d pictures are decoded", s, views_[s].hdr->max_ref, count[s]);
        }
        const auto tp0 = std::chrono::steady_clock::now();
        int rc = ensure_arena(a, L.bytes);
        if (rc != MOBI_OK) return rc;
        const auto tp1 = std::chrono::steady_clock::now();
        DevJob* jobs = reinterpret_cast<DevJob*>(a.h + L.jobs_off);
        const uint8_t** pics = reinterpret_cast<const uint8_t**>(a.h + L.pics_off);
        pool_.run(L.n_jobs, [&](int j) {
            const int s = L.job_stream[j];
            const mobi_packed_frame& f = views_[s];
            const mobi_frame_hdr& h = *f.hdr;
            const Off& o = off[j];
            std::memcpy(a.h + o.hdr, &h, sizeof h);
            if (h.n_mb) std::memcpy(a.h + o.mbs, f.mbs, sizeof(mobi_mb) * h.n_mb);
            if (h.n_parts) std::memcpy(a.h + o.parts, f.parts, sizeof(mobi_part) * h.n_parts);
            if (h.n_ops) std::memcpy(a.h + o.ops, f.ops, sizeof(mobi_op) * h.n_ops);
            if (h.n_coefs) std::memcpy(a.h + o.coefs, f.coefs, sizeof(mobi_coef) * h.n_coefs);
            if (h.n_intra) std::memcpy(a.h + o.intra, f.intra_list, sizeof(uint32_t) * h.n_intra);
            DevJob& J = jobs[j];
            std::memset(&J, 0, sizeof J);
            J.hdr = reinterpret_cast<const mobi_frame_hdr*>(a.d + o.hdr);
            J.mbs = reinterpret_cast<const mobi_mb*>(a.d + o.mbs);
            J.parts = reinterpret_cast<const mobi_part*>(a.d + o.parts);
            J.ops = reinterpret_cast<const mobi_op*>(a.d + o.ops);
            J.coefs = reinterpret_cast<const mobi_coef*>(a.d + o.coefs);
            J.intra = reinterpret_cast<const uint32_t*>(a.d + o.intra);
            const int c = count[s];
            J.dst = picture(s, c);
            pics[j] = J.dst;
            for (int k = 1; k <= 5; k++) J.ref[k - 1] = k <= c ? picture(s, c - k) : nullptr;
            J.dst_pic = (uint32_t)(s * RING + c % RING);
            for (int k = 1; k <= 5; k++) J.ref_pic[k - 1] = k <= c ? (uint32_t)(s * RING + (c - k) % RING) : 0u;
            J.flags = flags_ + (size_t)s * n_mb_;
            J.n_mb = h.n_mb; J.n_intra = h.n_intra;
        });
        // Intra work lists, each in dependency-depth order (depth = 1 + the deepest intra neighbour the macroblock
        // waits for): tickets drawn in this order never wait for a ticket that has not been drawn yet, and
        // macroblocks of equal depth -- of all pictures -- are independent, so the wavefronts of all streams advance
        // together.  Intra macroblocks of P-pictures (shallow, many) and of I-pictures (deep chains, few) go to separate
        // lists: they are launched on different CUDA streams so that the I-picture chains overlap the inter kernel.
        const auto tp2 = std::chrono::steady_clock::now();
        IntraWork* work = reinterpret_cast<IntraWork*>(a.h + L.work_off);
        depth_.resize(L.n_jobs);
        pool_.run(L.n_jobs, [&](int j) {
            const mobi_packed_frame& f = views_[L.job_stream[j]];
            const mobi_frame_hdr& h = *f.hdr;
            std::vector<uint16_t>& dep = depth_[j];
            dep.assign((size_t)h.n_mb + (size_t)h.n_intra * 2, 0);  // [0,n_mb): depth by MB; then (depth, wait) per rank
            uint16_t* per_rank = dep.data() + h.n_mb;
            for (uint32_t r = 0; r < h.n_intra; r++) {
                const int m = (int)f.intra_list[r];
                const uint32_t bits = (f.mbs[m].info >> 24) & 15u;
                uint32_t wait = 0, d = 0;
                for (int b = 0; b < 4; b++) {
                    if (!((bits >> b) & 1u)) continue;
                    const int nb = b == 0 ? m - 1 : m - g_.mbw - 2 + b;
                    if (nb < 0 || nb >= m || (f.mbs[nb].info & 3u) != 1u) continue;
                    wait |= 1u << b;
                    if (dep[nb] > d) d = dep[nb];
                }
                dep[m] = (uint16_t)(d + 1);
                per_rank[2 * r] = (uint16_t)(d + 1);
                per_rank[2 * r + 1] = (uint16_t)wait;
            }
            // Ticket order.  Depth order (what is awaited first) is a valid order, but it leaves the deepest, mutually dependent
            // macroblocks for the end of the launch, where nothing else is left to hide their latency (ncu: the SMs idle a
            // quarter of k_intra's run).  HEIGHT order -- the longest chain of dependents hanging off a macroblock, highest
            // first -- is a valid order too (whoever is awaited has a greater height than whoever waits, so it holds an earlier
            // ticket) and ends the launch with the independent leaves.  dep[] is reused: height by macroblock.
            std::fill(dep.begin(), dep.begin() + h.n_mb, (uint16_t)0);
            for (uint32_t r = h.n_intra; r-- > 0;) {
                const int m = (int)f.intra_list[r];
                const uint32_t wait = per_rank[2 * r + 1];
                for (int b = 0; b < 4; b++) {
                    if (!((wait >> b) & 1u)) continue;
                    const int nb = b == 0 ? m - 1 : m - g_.mbw - 2 + b;
                    if (dep[nb] < dep[m] + 1) dep[nb] = (uint16_t)(dep[m] + 1);
                }
            }
            for (uint32_t r = 0; r < h.n_intra; r++) per_rank[2 * r] = dep[f.intra_list[r]];   // height; turned into a sort key below
        });
        const auto tp3 = std::chrono::steady_clock::now();
        pack_ms_[0] += std::chrono::duration<double, std::milli>(tp1 - tp0).count();   // waiting for the arena (the GPU still reads the step before last)
        pack_ms_[1] += std::chrono::duration<double, std::milli>(tp2 - tp1).count();   // copies into the pinned arena + job table (pool)
        pack_ms_[2] += std::chrono::duration<double, std::milli>(tp3 - tp2).count();   // dependency heights (pool)
        struct Tail { double& acc; std::chrono::steady_clock::time_point t; ~Tail() { acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); } } tail_timer{pack_ms_[3], tp3};   // serial sort
        {
            uint32_t maxd = 0;
            for (int j = 0; j < L.n_jobs; j++) {
                const mobi_frame_hdr& h = *views_[L.job_stream[j]].hdr;
                const uint16_t* per_rank = depth_[j].data() + h.n_mb;
                for (uint32_t r = 0; r < h.n_intra; r++) if (per_rank[2 * r] > maxd) maxd = per_rank[2 * r];
            }
            for (int j = 0; j < L.n_jobs; j++) {   // sort key: greatest height first
                const mobi_frame_hdr& h = *views_[L.job_stream[j]].hdr;
                uint16_t* per_rank = depth_[j].data() + h.n_mb;
                for (uint32_t r = 0; r < h.n_intra; r++) per_rank[2 * r] = (uint16_t)(maxd - per_rank[2 * r]);
            }
            // P-pictures' intra macroblocks: counting sort by that key.  I-pictures: raster order per picture (the row kernel
            // indexes work[work_base + m]), pictures one after the other behind the P list.
            // One CTA per I-picture suits the few I-pictures of a staggered step (they hide behind the inter kernel on a few
            // SMs).  A step with more I-pictures than SMs would run them in waves of one picture-latency each (measured: 1024
            // pictures, 7 waves, 2.0 ms); such a step is throughput-bound, not latency-bound, and its macroblocks go through
            // the dependency-ordered ticket list with everything else (all pictures' wavefronts advance together).
            int n_ipics = 0;
            for (int j = 0; j < L.n_jobs; j++) { const mobi_frame_hdr& h = *views_[L.job_stream[j]].hdr; if (h.n_intra == h.n_mb) n_ipics++; }
            const bool ipics_by_ticket = n_ipics > key_ticket_threshold_;
            bucket_.assign((size_t)maxd + 2, 0);
            uint32_t* cnt = bucket_.data();
            for (int j = 0; j < L.n_jobs; j++) {
                const mobi_frame_hdr& h = *views_[L.job_stream[j]].hdr;
                if (h.n_intra == h.n_mb && !ipics_by_ticket) { L.n_key_jobs++; continue; }
                L.n_work_p += h.n_intra;
                const uint16_t* per_rank = depth_[j].data() + h.n_mb;
                for (uint32_t r = 0; r < h.n_intra; r++) cnt[per_rank[2 * r]]++;
            }
            uint32_t at = 0;
            for (uint32_t dd = 0; dd <= maxd + 1; dd++) { const uint32_t c = cnt[dd]; cnt[dd] = at; at += c; }
            uint32_t key_at = L.n_work_p, n_key = 0;
            uint32_t* keytab = reinterpret_cast<uint32_t*>(a.h + L.key_off);
            for (int j = 0; j < L.n_jobs; j++) {
                const mobi_packed_frame& f = views_[L.job_stream[j]];
                const mobi_frame_hdr& h = *f.hdr;
                const bool key = h.n_intra == h.n_mb && !ipics_by_ticket;
                const uint16_t* per_rank = depth_[j].data() + h.n_mb;
                if (key) { keytab[2 * n_key] = (uint32_t)j; keytab[2 * n_key + 1] = key_at; n_key++; }
                for (uint32_t r = 0; r < h.n_intra; r++) {
                    const uint32_t m = f.intra_list[r];
                    const mobi_mb& mb = f.mbs[m];
                    IntraWork& w = work[key ? key_at + r : cnt[per_rank[2 * r]]++];
                    w.job = (uint32_t)j; w.mb = m; w.info = mb.info; w.first_op = mb.first_sub; w.first_coef = mb.first_coef;
                    w.wait = per_rank[2 * r + 1]; w.pad[0] = w.pad[1] = 0;
                }
                if (key) key_at += h.n_intra;
            }
        }
        return MOBI_OK;
    }
    void rebase(uint8_t* h, const StepLayout& L, const uint8_t* from, uint8_t* to) {
        DevJob* jobs = reinterpret_cast<DevJob*>(h + L.jobs_off);
        const ptrdiff_t d = to - from;
        for (int j = 0; j < L.n_jobs; j++) {
            DevJob& J = jobs[j];
            J.hdr = reinterpret_cast<const mobi_frame_hdr*>(reinterpret_cast<const uint8_t*>(J.hdr) + d);
            J.mbs = reinterpret_cast<const mobi_mb*>(reinterpret_cast<const uint8_t*>(J.mbs) + d);
            J.parts = reinterpret_cast<const mobi_part*>(reinterpret_cast<const uint8_t*>(J.parts) + d);
            J.ops = reinterpret_cast<const mobi_op*>(reinterpret_cast<const uint8_t*>(J.ops) + d);
            J.coefs = reinterpret_cast<const mobi_coef*>(reinterpret_cast<const uint8_t*>(J.coefs) + d);
            J.intra = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(J.intra) + d);
        }
    }
    int launch_step(const uint8_t* d, const StepLayout& L) {
        const DevJob* jobs = reinterpret_cast<const DevJob*>(d + L.jobs_off);
        const IntraWork* work = reinterpret_cast<const IntraWork*>(d + L.work_off);
        const uint32_t n_key = L.n_work - L.n_work_p;
        stamp_++;
        if (n_key) {
            // I-pictures: long dependency chains, little parallelism (about mbw/2 macroblocks per picture at a time).
            // One CTA per picture on the side stream, concurrent with this step's inter kernel.
            if (!ok(cudaEventRecord(fork_, stream_), "cudaEventRecord")) return MOBI_ERR_CUDA;
            if (!ok(cudaStreamWaitEvent(side_, fork_, 0), "cudaStreamWaitEvent")) return MOBI_ERR_CUDA;
            if (timing_) tick(2, side_);
            if (!ok(launch_intra_key(jobs, work, d + L.key_off, L.n_key_jobs, g_, ticket_ + 48, side_), "k_intra_key")) return MOBI_ERR_CUDA;
            if (timing_) tick(2, side_);
            stats_.launches++;
            if (!ok(cudaEventRecord(join_, side_), "cudaEventRecord")) return MOBI_ERR_CUDA;
            key_resident_ += (uint32_t)L.n_key_jobs;
            if (L.n_inter_jobs) {   // let the few large I-picture CTAs settle before the flood of small k_inter CTAs
                if (!ok(launch_gate(ticket_ + 48, key_resident_, stream_), "k_gate")) return MOBI_ERR_CUDA;
                stats_.launches++;
            }
        }
        if (L.n_inter_jobs) {
            // k_mc then k_res: with timing on, the event between them closes the first kernel's bracket and opens the second's
            cudaEvent_t mid[2] = {nullptr, nullptr};
            if (timing_) { tick(0, stream_); mid[0] = fresh_event(); mid[1] = fresh_event(); }
            if (!ok(launch_inter(jobs, L.n_jobs, g_, tm_, sm_count_, ticket_ + 16, inter_ticket_base_, stream_, timing_ ? mid : nullptr), "k_mc / k_res")) return MOBI_ERR_CUDA;
            if (timing_) { ev_[0].push_back(mid[0]); ev_[3].push_back(mid[1]); tick(3, stream_); }
            stats_.launches += 2;
        }
        if (L.n_work_p) {
            uint32_t warps = 0;
            if (timing_) tick(1, stream_);
            if (!ok(launch_intra(jobs, work, L.n_work_p, ticket_, ticket_base_, stamp_, g_, (uint32_t)sm_count_ * 32u, stream_, &warps), "k_intra(P)")) return MOBI_ERR_CUDA;
            if (timing_) tick(1, stream_);
            ticket_base_ += L.n_work_p + warps;
            stats_.launches++;
        }
        if (n_key && !ok(cudaStreamWaitEvent(stream_, join_, 0), "cudaStreamWaitEvent")) return MOBI_ERR_CUDA;
        stats_.frames += L.n_jobs; stats_.mbs += L.mbs; stats_.inter_mbs += L.inter_mbs; stats_.intra_mbs += L.intra_mbs;
        stats_.parts += L.parts; stats_.coefs += L.coefs; stats_.ops += L.ops; stats_.inter_coefs += L.inter_coefs;
        return MOBI_OK;
    }

    int validate(const mobi_packed_frame& f, int pictures) {
        PackedValidator v(g_, H_);
        const int rc = v.run(f, pictures);
        if (rc != MOBI_OK) err_ = v.err_;
        return rc;
    }

    uint32_t W_, H_;
    int ver_, dev_, N_;
    Geom g_;
    uint32_t n_mb_;
    size_t ysz_, pic_;
    int sm_count_ = 148;
    int key_ticket_threshold_ = 148;   // more I-pictures than this in one step: ticket list instead of one CTA per picture (MOBI_KEY_TICKETS overrides)
    Pool pool_;
    std::vector<std::unique_ptr<Parser>> parsers_;
    std::vector<ParsedFrame> frames_;
    std::vector<mobi_packed_frame> views_;
    std::vector<int> rc_;
    std::vector<int> count_, staged_count_;
    cudaStream_t stream_ = nullptr;
    uint8_t* ring_ = nullptr;
    uint32_t* flags_ = nullptr;
    uint32_t* ticket_ = nullptr;
    uint32_t ticket_base_ = 0, inter_ticket_base_[2] = {0, 0}, stamp_ = 0, key_resident_ = 0;   // ticket_[0]: work tickets of k_intra; ticket_[16], [32]: chunk tickets of k_mc, k_res; ticket_[48]: resident I-picture CTAs
    InterMaps tm_;
    cudaStream_t side_ = nullptr, copy_ = nullptr, conv_ = nullptr;
    cudaEvent_t conv_go_ = nullptr, conv_done_ = nullptr;
    cudaEvent_t fork_ = nullptr, join_ = nullptr;
    std::vector<std::vector<uint16_t>> depth_;
    std::vector<uint32_t> bucket_;
    Arena arena_[2];
    int cur_arena_ = 0;
    std::vector<Staged> staged_;
    uint8_t* out_d_ = nullptr;
    uint8_t* out_h_ = nullptr;
    size_t out_cap_ = 0;
    const uint8_t** ptr_d_ = nullptr;
    const uint8_t** ring_tab_d_ = nullptr;
    const uint8_t** ptr_h_ = nullptr;
    bool timing_ = false;
    std::vector<cudaEvent_t> ev_[5], ev_free_;
    OutSlot slot_[2];
    int slot_head_ = 0, slots_used_ = 0;
    mobi_batch_stats stats_{};
    double phase_ms_[4] = {0, 0, 0, 0};
    double pack_ms_[4] = {0, 0, 0, 0};   // inside `pack`: arena wait, copies + job table, dependency heights, serial sort (MOBI_PACK_TRACE=1 prints them)
    uint32_t quant_override_ = 0, yuv_override_ = 0;
    bool have_override_ = false;
    std::string err_;
};

}  // namespace mobi

// ---- C ABI ---------------------------------------------------------------------------------------
struct mobi_batch { mobi::Batch b; mobi_batch(uint32_t w, uint32_t h, int v, int dev, int n, int t) : b(w, h, v, dev, n, t) {} };
struct mobi_decoder { mobi::Batch b; mobi_decoder(uint32_t w, uint32_t h, int v, int dev) : b(w, h, v, dev, 1, 1) {} };

namespace {
int check_geometry(uint32_t w, uint32_t h, int version) {
    if (w == 0 || h == 0 || (w & 15) || (h & 15) || w > 1024 || h > 1024) return MOBI_ERR_ARG;
    if (version == MOBI_VXDS) return MOBI_ERR_UNSUPPORTED;
    if (version != MOBI_MODSDS && version != MOBI_MOFLEX3DS) return MOBI_ERR_ARG;
    return MOBI_OK;
}
}  // namespace

extern "C" {

int mobi_packed_validate(uint32_t width, uint32_t height, int version, const mobi_packed_frame* f, int pictures, char* err, size_t err_len) {
    if (err && err_len) err[0] = 0;
    int rc = check_geometry(width, height, version);
    if (rc != MOBI_OK) return rc;
    if (!f || !f->hdr) return MOBI_ERR_ARG;
    mobi::Geom g;
    g.W = (int)width; g.H = (int)height; g.S = mobi::stride_for(width); g.version = version;
    g.log2S = g.S == 256 ? 8 : g.S == 512 ? 9 : 10;
    g.mbw = (int)width / 16; g.mbh = (int)height / 16;
    try {
        mobi::PackedValidator v(g, height);
        rc = v.run(*f, pictures);
        if (rc != MOBI_OK && err && err_len) { strncpy(err, v.err_.c_str(), err_len - 1); err[err_len - 1] = 0; }
        return rc;
    } catch (...) { return MOBI_ERR_NOMEM; }
}

int mobicuda_selftest_bgra(int device, unsigned long long* mismatches) {
    if (!mismatches) return MOBI_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return MOBI_ERR_CUDA;
    return mobi::selftest_bgra(mismatches) == cudaSuccess ? MOBI_OK : MOBI_ERR_CUDA;
}

int mobi_create(uint32_t width, uint32_t height, int version, int device, mobi_t** out) {
    if (!out) return MOBI_ERR_ARG;
    *out = nullptr;
    int rc = check_geometry(width, height, version);
    if (rc != MOBI_OK) return rc;
    mobi_decoder* d;
    try { d = new mobi_decoder(width, height, version, device); } catch (...) { return MOBI_ERR_NOMEM; }
    rc = d->b.init();
    if (rc != MOBI_OK) { fprintf(stderr, "mobicuda: %s\n", d->b.error()); delete d; return rc; }
    *out = d;
    return MOBI_OK;
}
void mobi_destroy(mobi_t* d) { delete d; }

int mobi_decode_frame(mobi_t* d, const uint8_t* data, int len, int* offset_inout) {
    if (!d) return MOBI_ERR_ARG;
    try { int st = 0; return d->b.decode(&data, &len, offset_inout, &st); } catch (...) { return d->b.set_err(MOBI_ERR_NOMEM, "out of host memory"); }
}
int mobi_submit_packed(mobi_t* d, const mobi_packed_frame* f) {
    if (!d) return MOBI_ERR_ARG;
    try { return d->b.submit(0, f); } catch (...) { return d->b.set_err(MOBI_ERR_NOMEM, "out of host memory"); }
}
int mobi_read_planes_strided(mobi_t* d, uint8_t* y, uint8_t* uv) { return d ? d->b.read_strided(0, y, uv) : MOBI_ERR_ARG; }
int mobi_read_yuv(mobi_t* d, uint8_t* y, uint8_t* u, uint8_t* v) { return d ? d->b.read_yuv_one(0, y, u, v) : MOBI_ERR_ARG; }
int mobi_read_bgra(mobi_t* d, uint8_t* dst, int dst_stride) { return d ? d->b.read_bgra_one(0, dst, dst_stride) : MOBI_ERR_ARG; }
int mobi_get_state(const mobi_t* d, uint32_t* quantizer, uint32_t* yuv_format, int* stride) {
    if (!d) return MOBI_ERR_ARG;
    d->b.get_state(0, quantizer, yuv_format, stride);
    return MOBI_OK;
}
const char* mobi_last_error(const mobi_t* d) { return d ? d->b.error() : "null decoder"; }

int mobi_batch_create(uint32_t width, uint32_t height, int version, int device, int n_streams, int n_threads, mobi_batch_t** out) {
    if (!out) return MOBI_ERR_ARG;
    *out = nullptr;
    int rc = check_geometry(width, height, version);
    if (rc != MOBI_OK) return rc;
    if (n_streams < 1 || n_streams > 65535) return MOBI_ERR_ARG;
    if (n_threads <= 0) { n_threads = (int)std::thread::hardware_concurrency(); if (n_threads < 1) n_threads = 1; }
    if (n_threads > n_streams) n_threads = n_streams;
    mobi_batch* b;
    try { b = new mobi_batch(width, height, version, device, n_streams, n_threads); } catch (...) { return MOBI_ERR_NOMEM; }
    rc = b->b.init();
    if (rc != MOBI_OK) { fprintf(stderr, "mobicuda: %s\n", b->b.error()); delete b; return rc; }
    *out = b;
    return MOBI_OK;
}
void mobi_batch_destroy(mobi_batch_t* b) { delete b; }
int mobi_batch_decode(mobi_batch_t* b, const uint8_t* const* data, const int* len, int* offset_inout, int* status) {
    if (!b) return MOBI_ERR_ARG;
    try { return b->b.decode(data, len, offset_inout, status); } catch (...) { return b->b.set_err(MOBI_ERR_NOMEM, "out of host memory"); }
}
int mobi_batch_submit(mobi_batch_t* b, const uint8_t* const* data, const int* len, int* offset_inout, int* status, int format) {
    if (!b) return MOBI_ERR_ARG;
    try { return b->b.submit_async(data, len, offset_inout, status, format); } catch (...) { return b->b.set_err(MOBI_ERR_NOMEM, "out of host memory"); }
}
int mobi_batch_fetch(mobi_batch_t* b, uint8_t* dst, const uint8_t** view, size_t* bytes) { return b ? b->b.fetch(dst, view, bytes) : MOBI_ERR_ARG; }
int mobi_batch_read_yuv(mobi_batch_t* b, uint8_t* dst) { return b ? b->b.read_yuv_all(dst) : MOBI_ERR_ARG; }
int mobi_batch_read_planes_strided(mobi_batch_t* b, int stream, uint8_t* y, uint8_t* uv) { return b ? b->b.read_strided(stream, y, uv) : MOBI_ERR_ARG; }
int mobi_batch_read_bgra(mobi_batch_t* b, int stream, uint8_t* dst, int dst_stride) { return b ? b->b.read_bgra_one(stream, dst, dst_stride) : MOBI_ERR_ARG; }
int mobi_batch_read_bgra_all(mobi_batch_t* b, uint8_t* dst) { return b ? b->b.bgra_all(dst) : MOBI_ERR_ARG; }
const char* mobi_batch_last_error(const mobi_batch_t* b) { return b ? b->b.error() : "null batch"; }
int mobi_batch_stage(mobi_batch_t* b, const uint8_t* const* data, const int* len, int* offset_inout) {
    if (!b) return MOBI_ERR_ARG;
    try { return b->b.stage(data, len, offset_inout); } catch (...) { return b->b.set_err(MOBI_ERR_NOMEM, "out of host memory"); }
}
int mobi_batch_replay(mobi_batch_t* b, int first, int count) { return b ? b->b.replay(first, count, 0) : MOBI_ERR_ARG; }
int mobi_batch_replay_convert(mobi_batch_t* b, int first, int count, int format) { return b ? b->b.replay(first, count, format) : MOBI_ERR_ARG; }
int mobi_batch_staged_steps(const mobi_batch_t* b) { return b ? b->b.staged_steps() : 0; }
void mobi_batch_clear_staged(mobi_batch_t* b) { if (b) b->b.clear_staged(); }
int mobi_batch_reset(mobi_batch_t* b) { return b ? b->b.reset(false) : MOBI_ERR_ARG; }
int mobi_batch_reset_streams(mobi_batch_t* b) { return b ? b->b.reset(true) : MOBI_ERR_ARG; }
int mobi_batch_sync(mobi_batch_t* b) { return b ? b->b.sync() : MOBI_ERR_ARG; }
void* mobi_batch_cuda_stream(mobi_batch_t* b) { return b ? b->b.cuda_stream() : nullptr; }
int mobi_batch_get_stats(const mobi_batch_t* b, mobi_batch_stats* st) {
    if (!b || !st) return MOBI_ERR_ARG;
    *st = b->b.stats();
    return MOBI_OK;
}
void mobi_batch_clear_stats(mobi_batch_t* b) { if (b) b->b.clear_stats(); }
int mobi_batch_get_phase_times(const mobi_batch_t* b, double ms[4]) {
    if (!b || !ms) return MOBI_ERR_ARG;
    b->b.phase_times(ms);
    return MOBI_OK;
}
int mobi_batch_set_kernel_timing(mobi_batch_t* b, int enabled) {
    if (!b) return MOBI_ERR_ARG;
    b->b.set_timing(enabled != 0);
    return MOBI_OK;
}
int mobi_batch_get_kernel_times(mobi_batch_t* b, double ms[5], uint64_t launches[5]) {
    return b ? b->b.kernel_times(ms, launches) : MOBI_ERR_ARG;
}

}  // extern "C"
