// pda_host_stage.h -- helpers for the *_host entry points: a grow-only device arena per
// device (so a batch-of-one shim call does not pay cudaMalloc/cudaFree every time) and
// checked async copies.  One *_host call at a time (g_hostMu).
#ifndef PDA_HOST_STAGE_H
#define PDA_HOST_STAGE_H

#include "pda_internal.h"

#include <cstring>
#include <mutex>
#include <vector>

namespace pda {

extern std::mutex g_hostMu;
struct DevArena { int device; unsigned char* base; size_t cap; };
extern std::vector<DevArena> g_arenas;

class Stage {
public:
    explicit Stage(int device) : device_(device), used_(0), base_(nullptr) {}
    // first pass: reserve() everything; then commit(); then at<T>(offset)
    size_t reserve(size_t bytes) { size_t o = used_; used_ += (bytes + 255) / 256 * 256; return o; }
    int commit() {
        for (DevArena& a : g_arenas)
            if (a.device == device_) {
                if (a.cap < used_) {
                    if (a.base) cudaFree(a.base);
                    a.base = nullptr; a.cap = 0;
                    size_t want = used_ + used_ / 4;
                    cudaError_t e = cudaMalloc(&a.base, want);
                    if (e != cudaSuccess) { (void)cudaGetLastError(); want = used_; e = cudaMalloc(&a.base, want); }
                    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(host staging arena)");
                    a.cap = want;
                }
                base_ = a.base;
                return PDA_OK;
            }
        DevArena a = {device_, nullptr, 0};
        g_arenas.push_back(a);
        return commit();
    }
    template <class T> T* at(size_t off) const { return reinterpret_cast<T*>(base_ + off); }
    size_t used() const { return used_; }
private:
    int device_; size_t used_; unsigned char* base_;
};

// three non-blocking streams per device for the chunked host path: copy in / run / copy out
struct HostStreams { int device; cudaStream_t in, run, out; };
extern std::vector<HostStreams> g_hostStreams;
inline int host_streams(int device, HostStreams** out) {
    for (HostStreams& h : g_hostStreams)
        if (h.device == device) { *out = &h; return PDA_OK; }
    HostStreams h;
    h.device = device;
    PDA_CUDA_TRY(cudaStreamCreateWithFlags(&h.in, cudaStreamNonBlocking));
    PDA_CUDA_TRY(cudaStreamCreateWithFlags(&h.run, cudaStreamNonBlocking));
    PDA_CUDA_TRY(cudaStreamCreateWithFlags(&h.out, cudaStreamNonBlocking));
    g_hostStreams.push_back(h);
    *out = &g_hostStreams.back();
    return PDA_OK;
}

inline int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        return fail(PDA_ERR_CUDA, "no CUDA device available (libpda_b200 has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(PDA_ERR_INVALID, "device %d out of range (have %d)", device, n);
    PDA_CUDA_TRY(cudaSetDevice(device));
    return PDA_OK;
}

template <class T> inline int h2d(T* dst, const T* src, size_t n, cudaStream_t s) {
    if (n == 0) return PDA_OK;
    PDA_CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return PDA_OK;
}
template <class T> inline int d2h(T* dst, const T* src, size_t n, cudaStream_t s) {
    if (n == 0 || dst == nullptr) return PDA_OK;
    PDA_CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, s));
    return PDA_OK;
}
#define PDA_TRY(expr) do { int _rc = (expr); if (_rc != PDA_OK) return _rc; } while (0)

// Small calls (a batch of one from the C++ shims, a handful of window frames) are dominated by the fixed cost of each
// cudaMemcpy from pageable memory (~10 us apiece, a dozen per call).  PackedIO lays all inputs out contiguously in the
// device arena, then all outputs, mirrors that layout in ONE pinned host buffer, and moves each side with a single
// asynchronous copy: reserve in()s, then out()s, then (after Stage::commit) upload() -> launches -> download().
struct PinnedBuf { unsigned char* p; size_t cap; };
extern PinnedBuf g_pinned;
constexpr size_t PDA_PACKED_LIMIT = 8u << 20;  // bytes of inputs + outputs below which a *_host call takes this path

class PackedIO {
public:
    explicit PackedIO(Stage& st) : st_(st), begin_(st.used()), inEnd_(st.used()), end_(st.used()) {}
    size_t in(const void* src, size_t bytes) {
        const size_t o = st_.reserve(bytes);
        if (src && bytes) items_.push_back({o, const_cast<void*>(src), bytes, true});
        inEnd_ = end_ = st_.used();
        return o;
    }
    size_t out(void* dst, size_t bytes) {
        const size_t o = st_.reserve(bytes);
        if (dst && bytes) items_.push_back({o, dst, bytes, false});
        end_ = st_.used();
        return o;
    }
    size_t bytes() const { return end_ - begin_; }
    int upload(cudaStream_t s) {
        const size_t need = end_ - begin_;
        if (g_pinned.cap < need) {
            if (g_pinned.p) cudaFreeHost(g_pinned.p);
            g_pinned.p = nullptr; g_pinned.cap = 0;
            const size_t want = need + need / 2 + 4096;
            PDA_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&g_pinned.p), want, cudaHostAllocDefault));
            g_pinned.cap = want;
        }
        for (const Item& it : items_)
            if (it.input) memcpy(g_pinned.p + (it.off - begin_), it.host, it.bytes);
        if (inEnd_ > begin_)
            PDA_CUDA_TRY(cudaMemcpyAsync(st_.at<unsigned char>(begin_), g_pinned.p, inEnd_ - begin_, cudaMemcpyHostToDevice, s));
        return PDA_OK;
    }
    int download(cudaStream_t s) {
        if (end_ > inEnd_)
            PDA_CUDA_TRY(cudaMemcpyAsync(g_pinned.p + (inEnd_ - begin_), st_.at<unsigned char>(inEnd_), end_ - inEnd_,
                                         cudaMemcpyDeviceToHost, s));
        PDA_CUDA_TRY(cudaStreamSynchronize(s));
        for (const Item& it : items_)
            if (!it.input) memcpy(it.host, g_pinned.p + (it.off - begin_), it.bytes);
        return PDA_OK;
    }
private:
    struct Item { size_t off; void* host; size_t bytes; bool input; };
    Stage& st_;
    size_t begin_, inEnd_, end_;
    std::vector<Item> items_;
};

}  // namespace pda
#endif
