// GPU-resident sigma builder: device copies of the replacement lists, integral repack, launch
// planning and the C ABI of include/forte2_b200.h (CISigmaBuilder half).
//
// Reference behaviour followed: forte2/ci/ci_sigma_builder.cc:29-222 (constructor, set_algorithm,
// set_memory, set_Hamiltonian validation, Hamiltonian = H0 + one-electron + same-spin + opposite
// spin, avg_build_time), forte2/determinant/slater_rules.cc:86-121 (determinant energies).
#include <algorithm>
#include <condition_variable>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "handles.hpp"
#include "kernels.cuh"
#include "lists_device.hpp"
#include "nvtx_ranges.hpp"
#include "rdm_kernels.cuh"
#include "same_spin_staged.cuh"

namespace f2 {

#define F2_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t err__ = (call);                                                                \
        if (err__ != cudaSuccess)                                                                  \
            throw CudaError(std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " +     \
                            __FILE__ + ":" + std::to_string(__LINE__));                            \
    } while (0)

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        if (count == n && p)
            return;
        release();
        if (count) {
            F2_CUDA(cudaMalloc(&p, count * sizeof(T)));
            n = count;
        }
    }
    void upload(const std::vector<T>& v) {
        alloc(std::max<size_t>(v.size(), 1));
        if (!v.empty())
            F2_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
};

struct DevStrings {
    int device = 0;
    DevBuf<uint32_t> a_cls, a_loc, b_cls, b_loc, as_ptr, nA_cls, nB_cls, a_off, b_off, zeros;
    DevBuf<uint64_t> as_ent, bs_ent, a_str, b_str;
    DevBuf<uint32_t> h2_ptr[2], h2_clist[2];
    DevBuf<uint64_t> h2_ent[2];
    DevBuf<int64_t> blk_off;  // [nclsA*nclsB]
    DevBuf<int64_t> blk_offT; // [nclsB*nclsA] same offsets indexed (cB, cA)
    bool al16 = false;        // every block offset and leading dimension is even
};

void destroy_dev_strings(DevStrings* d) {
    if (!d)
        return;
    int cur = 0;
    const bool have = cudaGetDevice(&cur) == cudaSuccess;
    cudaSetDevice(d->device); // the buffers are freed on the device that owns them
    delete d;
    if (have)
        cudaSetDevice(cur);
}

// the device copy of the lists for `device` (the caller has made it current); created once per device
static DevStrings* device_strings(f2_strings* S, int device) {
    std::lock_guard<std::mutex> lock(S->dev_mutex);
    auto it = S->dev.find(device);
    if (it != S->dev.end())
        return it->second;
    const CIStringsHost& h = *S->host;
    // the lists are made on this GPU unless the host has made them already (F2_HOST_LISTS: the host builder always)
    if (!h.lists_built() && std::getenv("F2_HOST_LISTS") == nullptr)
        build_lists_on_device(h);
    const DeviceLists& L = h.device_lists();
    auto d = std::make_unique<DevStrings>();
    d->device = device;
    d->a_cls.upload(L.a_cls);
    d->a_loc.upload(L.a_loc);
    d->b_cls.upload(L.b_cls);
    d->b_loc.upload(L.b_loc);
    d->as_ptr.upload(L.as_ptr);
    d->as_ent.upload(L.as_ent);
    d->bs_ent.upload(L.bs_ent);
    for (int s = 0; s < 2; ++s) {
        d->h2_ptr[s].upload(L.h2_ptr[s]);
        d->h2_ent[s].upload(L.h2_ent[s]);
        d->h2_clist[s].upload(L.h2_color_list[s]);
    }
    const auto& A = h.space(0, K_STR_N);
    const auto& B = h.space(1, K_STR_N);
    std::vector<uint32_t> nA(A.nclasses()), nB(B.nclasses()), aoff(A.nclasses()), boff(B.nclasses());
    std::vector<uint64_t> astr, bstr;
    for (size_t c = 0; c < A.nclasses(); ++c) {
        nA[c] = static_cast<uint32_t>(A.strpcls(c));
        aoff[c] = A.cls_off[c];
        astr.insert(astr.end(), A.cls[c].begin(), A.cls[c].end());
    }
    for (size_t c = 0; c < B.nclasses(); ++c) {
        nB[c] = static_cast<uint32_t>(B.strpcls(c));
        boff[c] = B.cls_off[c];
        bstr.insert(bstr.end(), B.cls[c].begin(), B.cls[c].end());
    }
    d->nA_cls.upload(nA);
    d->nB_cls.upload(nB);
    d->a_off.upload(aoff);
    d->b_off.upload(boff);
    d->a_str.upload(astr);
    d->b_str.upload(bstr);
    d->zeros.upload(std::vector<uint32_t>(std::max(A.nclasses(), B.nclasses()) + 1, 0u));
    std::vector<int64_t> off(A.nclasses() * B.nclasses(), -1), offT(A.nclasses() * B.nclasses(), -1);
    bool even = true;
    for (const auto& b : h.blocks) {
        off[b.cA * B.nclasses() + b.cB] = b.offset;
        offT[b.cB * A.nclasses() + b.cA] = b.offset;
        if (b.nA * b.nB == 0)
            continue;
        even = even && (b.offset % 2 == 0) && (b.nA % 2 == 0) && (b.nB % 2 == 0);
    }
    d->blk_off.upload(off);
    d->blk_offT.upload(offT);
    d->al16 = even;
    DevStrings* out = d.release();
    S->dev[device] = out;
    return out;
}

// one same-spin launch plan: which column groups exist, where each (row class, group) sits
struct SSPlan {
    int ngroups = 0, ntiles = 0;
    DevBuf<int64_t> xoff, yoff;
    DevBuf<uint32_t> g_ncols, g_ldx, g_ldy, tiles;
    bool al16 = false;
    // several determinant blocks: per column group the (N-2)-electron strings with a parent in the group
    DevBuf<uint32_t> gk_ptr, gk_list;
    uint32_t gk_max = 0; // longest group list
};

struct BetaPass {
    SSPlan plan;
    DevBuf<TrBlock> tr;
    std::vector<TrBlock> tr_host; // the same blocks: which sigma rows this pass completes
    int ntr = 0;
    uint32_t ntr_tiles = 0;
    int64_t elems = 0; // doubles used in each scratch buffer
};

// Everything that depends on which alpha rows (and, for the alpha same-spin term, which beta
// columns) this process owns.  rank 0 of world 1 is the single-GPU case.
struct Part {
    int rank = 0, world = 1;
    std::vector<int64_t> lo, hi;  // per alpha class: owned rows [lo, hi)
    std::vector<int64_t> cl, ch;  // per beta class: owned columns [cl, ch)
    std::vector<int64_t> loc_off; // per block: offset of the owned row slab in the local vector
    std::vector<int64_t> y_off;   // per block: offset of [nA][owned columns] in the exchange buffer
    int64_t local_size = 0, y_size = 0;
    DevBuf<int64_t> sblk_off;
    DevBuf<uint32_t> srow0;
    DevBuf<uint32_t> items;
    int nitems = 0;
    int64_t t3_kcols = 0;
    SSPlan alpha_plan;
    // rows form (sharded build without a full copy of C): alpha same-spin plan whose operand is the
    // exchanged column slab [nIa][owned columns], and the reduce-scatter row layout of the gather-form
    // opposite-spin kernel
    SSPlan alpha_cols_plan;
    DevBuf<uint32_t> rowpad;
    int64_t slot_rows = 0;  // single block: rows per reduce-scatter slot (max over ranks of the owned rows)
    int64_t slot_elems = 0; // doubles per reduce-scatter slot (max over ranks of the local vector length)
    // several blocks (t3gp_kernel): owned input strings, their local C offsets per beta class, and the offset
    // of every (alpha string, beta class) sigma segment in the slot layout
    DevBuf<uint32_t> jlist, crow_local, srow_slot;
    int njlist = 0;
    std::vector<std::unique_ptr<BetaPass>> beta_passes;
    bool beta_planned = false;
    DevBuf<CopyPiece> pieces; // owned row slabs: src = full layout, dst = local layout
    int npieces = 0;
    uint32_t piece_ctas = 0;
};

} // namespace f2

using namespace f2;

struct f2_dist; // NCCL state of the sharded build (dist_nccl.inc)
void destroy_dist(f2_dist* d);

struct f2_sigma {
    f2_strings* S = nullptr;
    DevStrings* D = nullptr;
    int device = 0, norb = 0, log_level = 3;
    double E = 0.0;
    std::string algorithm = "Knowles-Handy"; // C++ member default of the reference
    size_t memory_size = size_t(1) << 30;    // bytes; reference default 1 GiB
    // integral tables
    DevBuf<double> Vx, W[2], ea, eb, ja;
    DevBuf<double> kx, psame; // exchange integrals <pq|qp> and P_same(N_open): the spin-averaged diagonal
    DevBuf<uint16_t> rowmap;
    DevBuf<uint32_t> tgt; // [M][nbs] beta replacement table
    int tgt_pair = 0;
    int t3_ns = 2;
    // gather-form opposite-spin kernel (t3g_kernel): single determinant block whose rows fit in smem
    bool t3g = false;
    bool t3g_long = false; // the row of C does not fit in shared memory: part staged, the rest gathered from L2
    int t3g_nsm = 0;       // staged columns
    int t3g_nm16 = 0, t3g_has8 = 0, t3g_nx = 0, t3g_astride = 0, t3g_mpad = 0;
    int64_t t3g_gld = 0;
    size_t t3g_smem = 0;
    DevBuf<uint32_t> gsrc;
    DevBuf<double> VxT;
    // gather form over several beta classes (t3gp_kernel): symmetry / GAS blocks, one GPU
    bool t3gm = false;
    DevBuf<uint32_t> gm_gsrc, gm_gpos, gm_crow;
    DevBuf<uint16_t> gm_ml_ent;
    DevBuf<int32_t> gm_ml_cnt;
    int gm_mpad = 0, gm_nt16 = 0, gm_astride = 0;
    int64_t gm_gld = 0;
    size_t gm_smem = 0;
    bool t3_tail = false;  // last <= 8 integral rows spread over the warps
    bool fold_ok = false;  // alpha same-spin term can ride in the opposite-spin kernel (one GPU)
    DevBuf<double> kk2, Wfold; // 2 kk_ij / na per integral row; beta W with h - 2 kk folded
    int M = 0, P[2] = {0, 0};
    bool compressed = true;
    // opposite-spin kernel shape
    int t3_ks = 0, t3_warps = 0, t3_msplit = 1;
    // shape-generic kernels (t3_generic_kernel / ss_generic_kernel) for lists longer than the tensor-core kernels hold
    bool t3_generic = false;
    int t3_gen_rows = 0, t3_gen_tw = 0;
    bool ss_generic[2] = {false, false};
    int ss_gen_rows[2] = {0, 0}, ss_gen_tw[2] = {0, 0};
    size_t t3_smem = 0;
    // row/column ownership plans: the single-GPU one and the last multi-GPU one
    std::unique_ptr<Part> single, shard;
    std::unique_ptr<f2_dist, void (*)(f2_dist*)> dist{nullptr, destroy_dist};
    // host path: the single-GPU partition with the beta same-spin term cut into about eight row slabs,
    // so that finished sigma rows leave for the host while the next slab is computed
    std::unique_ptr<Part> host_part;
    DevBuf<double> XT, YT;
    int ss_ks[2] = {0, 0};
    // hdiag
    DevBuf<HdiagBlock> hd_blocks;
    int hd_nblocks = 0;
    uint32_t hd_rows = 0;
    // timing
    // per-term timing events: a ring of sets so that the host never has to wait for the previous build
    // before launching the next one (it may run up to kEvSets - 1 builds ahead of the device)
    static constexpr int kEvSets = 4;
    cudaEvent_t evs[kEvSets][5] = {};
    cudaEvent_t* ev = evs[0];
    int ev_idx = 0;
    bool set_pending[kEvSets] = {false, false, false, false};
    bool t3_first[kEvSets] = {false, false, false, false}; // rows form: the opposite-spin term runs first
    double t_acc[4] = {0, 0, 0, 0};
    double last_ms[5] = {0, 0, 0, 0, 0};
    long builds = 0;
    mutable int64_t launches = 0; // kernels launched by the last build
    // t3gp_kernel: shape (k-steps per batch, row tiles per item) chosen by timing the candidates on the first builds
    struct T3gpTune {
        static constexpr int kCand = 3;   // instances that take part in the timing
        static constexpr int kShapes = 6; // instances that exist (F2_T3GP_SHAPE pins any of them)
        int choice = -1;   // index of the chosen candidate, -1 while trying
        int calls = 0;     // launches so far (the first one is not timed: first-use costs)
        int trial = 0;     // candidate of the launch whose events are pending
        bool pending = false;
        float ms[kCand] = {};
        cudaEvent_t ev[2] = {nullptr, nullptr};
    };
    mutable T3gpTune t3gp_tune;
    // reduced density matrices: hole lists of (spin, level) on the device (uploaded on first use) and the two
    // D buffers of the gather + GEMM formulation (csrc/rdm_kernels.cuh)
    struct RdmLists {
        DevBuf<uint32_t> ptr;
        DevBuf<uint64_t> ent;
        bool up = false;
    } rdm_lists[2][4];
    DevBuf<double> rdmD[2];
    DevBuf<uint32_t> rdm_gtab; // opposite-spin 2-RDM: source column and phase of every ((v y), Ib) (rdm_ab_kernel)
    int64_t rdm_gld = 0;
    int rdm_gtab_rows = 0;
    // host staging for f2_sigma_apply_host
    double *hC = nullptr, *hS = nullptr;
    DevBuf<double> dC, dS;
    cudaStream_t own_stream = nullptr, own_stream2 = nullptr, copy_stream = nullptr;
    cudaEvent_t t3_ev = nullptr;
    cudaEvent_t copy_ev = nullptr, pass_ev = nullptr;
    std::vector<cudaEvent_t> piece_ev;

    ~f2_sigma() {
        for (auto& e : piece_ev)
            cudaEventDestroy(e);
        for (auto& set : evs)
            for (auto& e : set)
                if (e)
                    cudaEventDestroy(e);
        if (hC)
            cudaFreeHost(hC);
        if (hS)
            cudaFreeHost(hS);
        if (own_stream)
            cudaStreamDestroy(own_stream);
        if (own_stream2)
            cudaStreamDestroy(own_stream2);
        if (t3_ev)
            cudaEventDestroy(t3_ev);
        for (cudaEvent_t e : t3gp_tune.ev)
            if (e)
                cudaEventDestroy(e);
        if (copy_stream)
            cudaStreamDestroy(copy_stream);
        if (copy_ev)
            cudaEventDestroy(copy_ev);
        if (pass_ev)
            cudaEventDestroy(pass_ev);
    }
};

namespace {

// host-to-host copy split over a few threads (one memcpy stream does not saturate the memory system)
// Host copies between the caller's (pageable) vectors and the pinned staging buffers: a persistent pool of threads (no
// thread creation per 64 MB piece) that writes with non-temporal stores (no read-for-ownership of the destination, no
// cache pollution).  The way home of sigma is bound by this copy (profiles/r2r_host_path.log): 39 GB/s with 12 threads
// and plain memcpy.
class CopyPool {
  public:
    static CopyPool& get() {
        static CopyPool pool;
        return pool;
    }
    void copy(double* dst, const double* src, size_t count) {
        if (count < (size_t(1) << 18) || workers_.empty()) {
            stream_copy(dst, src, count);
            return;
        }
        std::unique_lock<std::mutex> lock(m_);
        dst_ = dst;
        src_ = src;
        count_ = count;
        pending_ = static_cast<int>(workers_.size());
        ++generation_;
        cv_.notify_all();
        done_.wait(lock, [&] { return pending_ == 0; });
    }

  private:
    CopyPool() {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const unsigned n = std::min(hw, 16u);
        if (n > 1)
            for (unsigned t = 0; t < n; ++t)
                workers_.emplace_back([this, t, n] { run(t, n); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lock(m_);
            stop_ = true;
            ++generation_;
        }
        cv_.notify_all();
        for (auto& w : workers_)
            w.join();
    }
    static void stream_copy(double* dst, const double* src, size_t count) {
#if defined(__SSE2__)
        size_t i = 0;
        while (i < count && (reinterpret_cast<uintptr_t>(dst + i) & 15) != 0) {
            dst[i] = src[i];
            ++i;
        }
        for (; i + 8 <= count; i += 8) {
            const __m128d a = _mm_loadu_pd(src + i), b = _mm_loadu_pd(src + i + 2);
            const __m128d c = _mm_loadu_pd(src + i + 4), d = _mm_loadu_pd(src + i + 6);
            _mm_stream_pd(dst + i, a);
            _mm_stream_pd(dst + i + 2, b);
            _mm_stream_pd(dst + i + 4, c);
            _mm_stream_pd(dst + i + 6, d);
        }
        for (; i < count; ++i)
            dst[i] = src[i];
        _mm_sfence();
#else
        std::memcpy(dst, src, count * sizeof(double));
#endif
    }
    void run(unsigned t, unsigned n) {
        unsigned long seen = 0;
        for (;;) {
            double* dst;
            const double* src;
            size_t count;
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_)
                    return;
                dst = dst_;
                src = src_;
                count = count_;
            }
            const size_t per = ((count + n - 1) / n + 7) & ~size_t(7);
            const size_t lo = std::min(count, t * per), hi = std::min(count, lo + per);
            if (lo < hi)
                stream_copy(dst + lo, src + lo, hi - lo);
            {
                std::lock_guard<std::mutex> lock(m_);
                if (--pending_ == 0)
                    done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    double* dst_ = nullptr;
    const double* src_ = nullptr;
    size_t count_ = 0;
    int pending_ = 0;
    unsigned long generation_ = 0;
    bool stop_ = false;
};

void parallel_copy(double* dst, const double* src, size_t count) { CopyPool::get().copy(dst, src, count); }

int round_ks(int count) { // k-steps (of 4) rounded up to the instantiated multiples of 4
    const int ks = (count + 3) / 4;
    return std::max(4, ((ks + 3) / 4) * 4);
}

void build_integrals(f2_sigma& b, double E, const double* H, const double* V) {
    const CIStringsHost& h = *b.S->host;
    const int n = h.norb;
    b.E = E;
    auto v = [&](int p, int q, int r, int s) {
        return V[((static_cast<size_t>(p) * n + q) * n + r) * n + s];
    };
    // 8-fold (real orbital) symmetry allows the pair-compressed rows (i >= j)
    double vmax = 0.0, asym = 0.0;
    for (int k = 0; k < n; ++k)
        for (int i = 0; i < n; ++i)
            for (int l = 0; l < n; ++l)
                for (int j = 0; j < n; ++j) {
                    vmax = std::max(vmax, std::fabs(v(k, i, l, j)));
                    asym = std::max(asym, std::fabs(v(k, i, l, j) - v(k, j, l, i)));
                }
    b.compressed = asym <= 1e-12 * std::max(vmax, 1e-300);
    const int n2 = n * n;
    std::vector<uint16_t> rowmap(std::max(n2, 1));
    if (b.compressed) {
        b.M = n * (n + 1) / 2;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                const int hi = std::max(i, j), lo = std::min(i, j);
                rowmap[i * n + j] = static_cast<uint16_t>(hi * (hi + 1) / 2 + lo);
            }
    } else {
        b.M = n2;
        for (int i = 0; i < n2; ++i)
            rowmap[i] = static_cast<uint16_t>(i);
    }
    std::vector<double> Vx(static_cast<size_t>(std::max(b.M, 1)) * std::max(n2, 1), 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            if (b.compressed && j > i)
                continue;
            const size_t m = rowmap[i * n + j];
            for (int k = 0; k < n; ++k)
                for (int l = 0; l < n; ++l)
                    Vx[m * n2 + k * n + l] =
                        b.compressed ? 0.5 * (v(k, i, l, j) + v(k, j, l, i)) : v(k, i, l, j);
        }
    b.Vx.upload(Vx);
    b.rowmap.upload(rowmap);
    {
        std::vector<double> VxT(Vx.size());
        for (int m = 0; m < b.M; ++m)
            for (int kl = 0; kl < n2; ++kl)
                VxT[static_cast<size_t>(kl) * b.M + m] = Vx[static_cast<size_t>(m) * n2 + kl];
        b.VxT.upload(VxT);
    }

    // beta replacement table: (row of the integral matrix, source string Jb) -> target string Ib.
    // A row holds at most one replacement per source string (E_ij and E_ji exclude each other).
    {
        const DeviceLists& L = h.device_lists();
        const auto& B = h.space(1, K_STR_N);
        const size_t nbs = B.nstr();
        if (B.nclasses() > 2046)
            throw std::overflow_error("CISigmaBuilder: more than 2046 beta string classes");
        for (size_t c = 0; c < B.nclasses(); ++c)
            if (B.strpcls(c) > (size_t(1) << 20))
                throw std::overflow_error("CISigmaBuilder: more than 2^20 beta strings in one class");
        const DevStrings& D = *b.D;
        const int64_t nT = static_cast<int64_t>(b.M) * static_cast<int64_t>(nbs);
        auto grid_for = [](int64_t n) { return static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, int64_t(1) << 30)); };
        // (legacy stream throughout: every upload above is a blocking cudaMemcpy on it)
        b.tgt.alloc(static_cast<size_t>(std::max<int64_t>(nT, 1)));
        F2_CUDA(cudaMemset(b.tgt.p, 0xff, static_cast<size_t>(std::max<int64_t>(nT, 1)) * sizeof(uint32_t)));
        const int64_t nent = static_cast<int64_t>(nbs) * L.bs_width;
        if (nent > 0 && nT > 0)
            tgt_table_kernel<<<grid_for(nent), 256>>>(D.bs_ent.p, static_cast<int64_t>(nbs), L.bs_width, b.rowmap.p, b.tgt.p);
        // inverse table for the gather form: (row m, target Ib) -> source Jb with <Ib|E_m|Jb> = sign;
        // "no source" points at the zero slot behind the staged row (index nbs).  Single beta class only.
        b.gsrc.release();
        if (B.nclasses() == 1 && h.nclsA() == 1 && nbs < (size_t(1) << 20)) {
            const int Mpad = (b.M + 3) / 4 * 4;
            const size_t gld = (nbs + 15) / 16 * 16;
            const int64_t nG = static_cast<int64_t>(std::max<size_t>(size_t(Mpad) * gld, 1));
            b.gsrc.alloc(static_cast<size_t>(nG));
            fill_u32_kernel<<<148 * 8, 256>>>(b.gsrc.p, static_cast<uint32_t>(nbs), nG);
            // E_m Jb = sign Ib.  With pair-compressed rows E_ij and E_ji share row m; otherwise the
            // gather row of <Ib|E_ij|Jb> is the row of (i, j) itself
            if (nT > 0)
                gsrc_table_kernel<<<grid_for(nT), 256>>>(b.tgt.p, b.M, static_cast<int64_t>(nbs), static_cast<int64_t>(gld),
                                                         b.gsrc.p);
            b.t3g_mpad = Mpad;
            b.t3g_gld = static_cast<int64_t>(gld);
        }
        // the same for several beta classes (t3gp_kernel): table columns are the targets class by class,
        // every class padded to a multiple of 16; an entry names its source as class << 20 | address.  One
        // extra all-empty row.  Beside it, for every (alpha class of the input string, target beta class) the
        // integral rows whose source class has a determinant block with that alpha class.
        b.gm_gsrc.release();
        if (!b.gsrc.p && nbs > 0 && h.ndet < (int64_t(1) << 32) - 1 && std::getenv("F2_NO_T3GM") == nullptr) {
            const int Mpad = (b.M + 3) / 4 * 4;
            const size_t NCb = B.nclasses(), NCa = h.nclsA();
            std::vector<uint32_t> gpos(NCb);
            size_t gld = 0;
            for (size_t c = 0; c < NCb; ++c) {
                gpos[c] = static_cast<uint32_t>(gld);
                gld += (B.strpcls(c) + 15) / 16 * 16;
            }
            if (static_cast<size_t>(Mpad + 1) * gld * sizeof(uint32_t) <= (size_t(2) << 30) && Mpad < 65535) {
                const int64_t nG = static_cast<int64_t>(std::max<size_t>(size_t(Mpad + 1) * gld, 1));
                b.gm_gsrc.alloc(static_cast<size_t>(nG));
                fill_u32_kernel<<<148 * 8, 256>>>(b.gm_gsrc.p, 0x7fffffffu, nG);
                b.gm_gpos.upload(gpos);
                // [target class][m] -> source classes (a pair-compressed row serves E_ij and E_ji, which can come
                // from two different classes); -2 in slot 0 = more than two, keep the row unconditionally
                DevBuf<int> dsc;
                dsc.alloc(NCb * Mpad * 2);
                F2_CUDA(cudaMemset(dsc.p, 0xff, NCb * Mpad * 2 * sizeof(int))); // -1
                if (nT > 0)
                    gm_table_kernel<<<grid_for(nT), 256>>>(b.tgt.p, b.M, Mpad, static_cast<int64_t>(nbs), static_cast<int64_t>(gld),
                                                           b.gm_gpos.p, D.b_cls.p, D.b_loc.p, b.gm_gsrc.p, dsc.p);
                F2_CUDA(cudaGetLastError());
                std::vector<int> srccls(NCb * Mpad * 2, -1);
                F2_CUDA(cudaMemcpy(srccls.data(), dsc.p, srccls.size() * sizeof(int), cudaMemcpyDeviceToHost));
                std::vector<uint16_t> ml(NCa * NCb * Mpad, static_cast<uint16_t>(Mpad));
                std::vector<int32_t> mc(NCa * NCb, 0);
                for (size_t ca = 0; ca < NCa; ++ca)
                    for (size_t ct = 0; ct < NCb; ++ct) {
                        int n = 0;
                        for (int m = 0; m < b.M; ++m) {
                            const int* sc = &srccls[(ct * Mpad + m) * 2];
                            bool keep = sc[0] == -2;
                            for (int q = 0; q < 2 && !keep; ++q)
                                if (sc[q] >= 0) {
                                    const int64_t bi = h.block_of[ca * NCb + sc[q]];
                                    keep = bi >= 0 && h.blocks[bi].nA * h.blocks[bi].nB > 0;
                                }
                            if (keep)
                                ml[(ca * NCb + ct) * Mpad + n++] = static_cast<uint16_t>(m);
                        }
                        mc[ca * NCb + ct] = n;
                    }
                b.gm_ml_ent.upload(ml);
                b.gm_ml_cnt.upload(mc);
                b.gm_mpad = Mpad;
                b.gm_gld = static_cast<int64_t>(gld);
            }
        }
        F2_CUDA(cudaGetLastError());
        F2_CUDA(cudaDeviceSynchronize());
        bool pair = nbs % 2 == 0;
        for (size_t c = 0; c < B.nclasses(); ++c)
            pair = pair && (B.cls_off[c] % 2 == 0) && (B.strpcls(c) % 2 == 0);
        b.tgt_pair = pair ? 1 : 0;
    }

    // kk_ij = h_ij - 1/2 sum_m (im|mj): the one-electron operator of the E_ij E_kl form
    std::vector<double> kk(std::max(n2, 1), 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double x = H[i * n + j];
            for (int m = 0; m < n; ++m)
                x -= 0.5 * v(i, m, m, j);
            kk[i * n + j] = x;
        }
    {
        std::vector<double> kk2(std::max(b.M, 1), 0.0);
        if (h.na >= 1)
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j)
                    kk2[rowmap[i * n + j]] = 2.0 * kk[i * n + j] / h.na;
        b.kk2.upload(kk2);
    }

    // same-spin integrals: antisymmetrised pairs with the one-electron operator folded in.
    // variant 2 (beta only): one-electron matrix h - 2 kk, used when the opposite-spin kernel also
    // carries the alpha same-spin term (its beta scatter then leaves an extra 2 kk one-electron term)
    for (int variant = 0; variant < 3; ++variant) {
        const int spin = variant == 0 ? 0 : 1;
        const int ne = spin == 0 ? h.na : h.nb;
        std::vector<double> W, h1(std::max(n2, 1));
        for (int i = 0; i < n2; ++i)
            h1[i] = variant == 2 ? H[i] - 2.0 * kk[i] : H[i];
        const double* H = h1.data();
        if (ne == 1) {
            b.P[spin] = n;
            W.assign(H, H + static_cast<size_t>(n) * n);
        } else if (ne >= 2) {
            const int P = n * (n - 1) / 2;
            b.P[spin] = P;
            W.assign(static_cast<size_t>(P) * P, 0.0);
            const double f = 1.0 / (ne - 1);
            auto hh = [&](int p, int q) { return H[p * n + q]; };
            for (int p = 1; p < n; ++p)
                for (int q = 0; q < p; ++q)
                    for (int r = 1; r < n; ++r)
                        for (int s = 0; s < r; ++s) {
                            double w = v(p, q, r, s) - v(p, q, s, r);
                            if (q == s)
                                w += f * hh(p, r);
                            if (p == r)
                                w += f * hh(q, s);
                            if (q == r)
                                w -= f * hh(p, s);
                            if (p == s)
                                w -= f * hh(q, r);
                            W[static_cast<size_t>(p * (p - 1) / 2 + q) * P + r * (r - 1) / 2 + s] = w;
                        }
        } else {
            b.P[spin] = 0;
        }
        if (variant == 2)
            b.Wfold.upload(W);
        else
            b.W[spin].upload(W);
    }

    // determinant-energy tables
    std::vector<double> J(std::max(n2, 1)), K(std::max(n2, 1));
    for (int p = 0; p < n; ++p)
        for (int q = 0; q < n; ++q) {
            J[p * n + q] = v(p, q, p, q);
            K[p * n + q] = v(p, q, q, p);
        }
    auto string_energy = [&](const StringSpace& S, std::vector<double>& e, std::vector<double>* ja) {
        e.assign(std::max<size_t>(S.nstr(), 1), 0.0);
        if (ja)
            ja->assign(std::max<size_t>(S.nstr() * n, 1), 0.0);
        size_t g = 0;
        for (const auto& cls : S.cls)
            for (uint64_t w : cls) {
                double x = 0.0;
                for (int p = 0; p < n; ++p) {
                    if (!((w >> p) & 1))
                        continue;
                    x += H[p * n + p];
                    for (int q = 0; q < p; ++q)
                        if ((w >> q) & 1)
                            x += J[p * n + q] - K[p * n + q];
                    if (ja)
                        for (int q = 0; q < n; ++q)
                            (*ja)[g * n + q] += J[p * n + q];
                }
                e[g++] = x;
            }
    };
    {
        // P_same(No): of the C(No, nao) ways to give nao of the No singly occupied orbitals an alpha electron
        // (nao - nbo = na - nb), the fraction in which two given ones carry the same spin
        std::vector<double> ps(65, 0.0);
        auto binom = [](int n, int k) {
            if (k < 0 || k > n)
                return 0.0;
            double x = 1.0;
            for (int i = 1; i <= k; ++i)
                x = x * (n - k + i) / i;
            return x;
        };
        for (int no = 2; no <= 64; ++no) {
            const int twoms = h.na - h.nb;
            if ((no + twoms) % 2 != 0 || no + twoms < 0 || no - twoms < 0)
                continue;
            const int nao = (no + twoms) / 2;
            const double all = binom(no, nao);
            if (all > 0.0)
                ps[no] = (binom(no - 2, nao - 2) + binom(no - 2, nao)) / all;
        }
        b.kx.upload(K);
        b.psame.upload(ps);
    }
    std::vector<double> ea, eb, ja;
    string_energy(h.space(0, K_STR_N), ea, &ja);
    string_energy(h.space(1, K_STR_N), eb, nullptr);
    b.ea.upload(ea);
    b.eb.upload(eb);
    b.ja.upload(ja);
}

// widest strip of columns (a power of two, at most GEN_THREADS) whose staged rows fit in shared memory; 0 = none does
int generic_strip(int rows, int extra8) {
    for (int tw = GEN_THREADS; tw >= 4; tw >>= 1)
        if (gen_smem_bytes(rows, tw, extra8) <= size_t(200) * 1024)
            return tw;
    return 0;
}

void plan_t3_shape(f2_sigma& b) {
    const CIStringsHost& h = *b.S->host;
    const DeviceLists& L = h.device_lists();
    b.t3_ks = 0;
    if (h.na < 1 || h.nb < 1 || L.as_max == 0 || h.ndet == 0)
        return;
    b.t3_generic = false;
    auto go_generic = [&]() {
        b.t3_generic = true;
        b.t3_gen_rows = L.as_max;
        b.t3_gen_tw = generic_strip(L.as_max, static_cast<int>(h.nclsB()));
        if (b.t3_gen_tw == 0)
            throw std::runtime_error("CISigmaBuilder: " + std::to_string(L.as_max) + " single replacements per alpha "
                                     "string do not fit the opposite-spin kernels");
        b.t3_ks = 4; // non-zero: there is an opposite-spin term
        b.t3_msplit = 1;
        b.t3_warps = GEN_THREADS / 32;
        b.t3_tail = false;
        b.fold_ok = false;
        b.t3g = false;
        b.t3gm = false;
    };
    if (L.as_max > 128 || std::getenv("F2_GENERIC_KERNELS") != nullptr) {
        go_generic();
        return;
    }
    b.t3_ks = round_ks(L.as_max);
    const int MT = (b.M + 15) / 16;
    const int maxw = b.t3_ks <= 24 ? 12 : 10;
    const size_t nclsB = h.nclsB();
    const int msplit = (MT + maxw - 1) / maxw;
    b.t3_warps = (MT + msplit - 1) / msplit;
    b.t3_msplit = msplit;
    // a last tile of at most 8 rows is spread over the other warps when that balances the four SM
    // sub-partitions (4 or 8 warps left)
    b.t3_tail = msplit == 1 && (b.M - (MT - 1) * 16) <= 8 && (MT - 1 == 4 || MT - 1 == 8);
    if (b.t3_tail)
        b.t3_warps = MT - 1;
    const int rows = b.t3_warps * 16 + (b.t3_tail ? 8 : 0);
    b.t3_ns = t3_smem_bytes(b.t3_ks, 3, static_cast<int>(nclsB), rows) <= 227 * 1024 ? 3 : 2;
    b.t3_smem = t3_smem_bytes(b.t3_ks, b.t3_ns, static_cast<int>(nclsB), rows);
    // the alpha same-spin term can be folded in when every single replacement of every alpha string
    // stays in the string space (no GAS restriction), so that E_ij E_kl needs no outside intermediate
    b.fold_ok = L.as_ent.size() == static_cast<size_t>(h.nas) * (h.na * (h.norb - h.na) + h.na) &&
                std::getenv("F2_NO_FOLD") == nullptr;
    if (b.t3_smem > 227 * 1024) {
        go_generic();
        return;
    }
    // gather form: one determinant block, rows of C (and the A' block) fit in shared memory
    b.t3g = false;
    if (b.gsrc.p && b.D->al16 && h.nbs % 2 == 0 && b.M >= 33 && std::getenv("F2_NO_T3G") == nullptr) {
        // rows of A': the single replacements of a string, its na diagonal ones (E_kk) merged into one row
        const int kn = L.as_max - (h.na > 0 ? static_cast<int>(h.na) - 1 : 0);
        const int nm16 = kn / 16, rem = kn % 16;
        b.t3g_nm16 = nm16 + (rem > 8 ? 1 : 0);
        b.t3g_has8 = (rem > 0 && rem <= 8) ? 1 : 0;
        b.t3g_nx = 0;
        // one or two rows beyond the 16-row tiles go to the FMA pipe instead of an 8-row tile (t3g_kernel, NX):
        // instantiated for 4 and 5 row tiles (CAS(16,16), CAS(18,18) and their neighbours)
        // (measured, profiles/r2j_quick.log: a DFMA between DMMAs costs far more than its 2 issue cycles -- the gain is
        // 2.7 % at CAS(18,18), 0.9 % at CAS(16,16) and a loss with 3 row tiles, so 4 and 5 row tiles only)
        if (rem >= 1 && rem <= 2 && nm16 >= 4 && nm16 <= 5 && std::getenv("F2_T3G_NO_FMA_ROWS") == nullptr) {
            b.t3g_has8 = 0;
            b.t3g_nx = rem;
        }
        const int rows = b.t3g_nm16 * 16 + b.t3g_has8 * 8 + b.t3g_nx;
        int S = b.t3g_mpad;
        while (S % 16 != 4)
            ++S;
        b.t3g_astride = S;
        b.t3g_smem = t3g_smem_bytes(h.nbs, rows, S);
        b.t3g_long = false;
        b.t3g_nsm = static_cast<int>(h.nbs);
        const bool shape_ok = b.t3g_nm16 >= 1 && b.t3g_nm16 <= 6 && !(b.t3g_nm16 == 6 && b.t3g_has8);
        if (shape_ok && b.t3g_smem > 227 * 1024 - 2048 && std::getenv("F2_NO_T3G_LONG") == nullptr) {
            // long rows: stage as many columns as fit beside the A' block, gather the rest from L2
            const int64_t fixed = static_cast<int64_t>(t3g_smem_bytes(0, rows, S));
            int64_t nsm = (int64_t(227) * 1024 - 2048 - fixed) / 8; // 2 KB left for the kernel's static arrays
            nsm -= nsm % 2;
            if (const char* e = std::getenv("F2_T3G_NSM")) // tuning / test hook
                nsm = std::min<int64_t>(nsm, std::max<int64_t>(0, std::atoll(e)) & ~int64_t(1));
            if (nsm >= 0) {
                b.t3g_long = true;
                b.t3g_nsm = static_cast<int>(std::min<int64_t>(nsm, h.nbs));
                b.t3g_smem = t3g_smem_bytes(b.t3g_nsm, rows, S);
            }
        } else if (shape_ok && std::getenv("F2_T3G_NSM") != nullptr) {
            // test hook: exercise the long-row path on a small space
            const int64_t nsm = std::min<int64_t>(h.nbs, std::max<int64_t>(0, std::atoll(std::getenv("F2_T3G_NSM"))) & ~int64_t(1));
            b.t3g_long = true;
            b.t3g_nsm = static_cast<int>(nsm);
            b.t3g_smem = t3g_smem_bytes(b.t3g_nsm, rows, S);
        }
        b.t3g = shape_ok && b.t3g_smem <= 227 * 1024 - 2048;
    }
    // gather form over several beta classes: the A' block and the per-class row / k lists must fit
    b.t3gm = false;
    if (!b.t3g && b.gm_gsrc.p) {
        b.gm_nt16 = (L.as_max + 15) / 16;
        const int rows = b.gm_nt16 * 16;
        // row stride of A': 4 or 12 (mod 16) doubles puts the rows g = 0..3 of a half-warp's 64-bit loads on
        // disjoint banks; the finer step keeps 3 GAS spaces x 20 orbitals (M = 210, 25 classes) within 227 KB
        int S = b.gm_mpad + 1;
        while (S % 8 != 4)
            ++S;
        b.gm_astride = S;
        b.gm_smem = t3gp_smem_bytes(rows, S, static_cast<int>(nclsB), b.gm_mpad);
        b.t3gm = b.gm_nt16 >= 1 && b.gm_nt16 <= 8 && b.gm_smem <= 227 * 1024 - 2048;
        if (b.t3gm && !b.gm_crow.p) {
            // offset of C[Ja, class c, 0] for every alpha string and beta class (the CI-vector layout)
            const auto& A = h.space(0, K_STR_N);
            std::vector<uint32_t> crow(static_cast<size_t>(h.nas) * nclsB, kNoRow);
            for (size_t cA = 0; cA < A.nclasses(); ++cA)
                for (size_t cB = 0; cB < nclsB; ++cB) {
                    const int64_t bi = h.block_of[cA * nclsB + cB];
                    if (bi < 0 || h.blocks[bi].nA * h.blocks[bi].nB == 0)
                        continue;
                    const auto& blk = h.blocks[bi];
                    for (int64_t i = 0; i < blk.nA; ++i)
                        crow[(A.cls_off[cA] + i) * nclsB + cB] = static_cast<uint32_t>(blk.offset + i * blk.nB);
                }
            b.gm_crow.upload(crow);
        }
    }
}

void fill_tiles(std::vector<uint32_t>& tiles, const std::vector<uint32_t>& ncols) {
    tiles.clear();
    for (size_t g = 0; g < ncols.size(); ++g)
        for (uint32_t j0 = 0; j0 < ncols[g]; j0 += SSG_TW) {
            tiles.push_back(static_cast<uint32_t>(g));
            tiles.push_back(j0);
        }
}

// Spaces with several determinant blocks: most (N-2)-electron strings K have no parent in the blocks of a given
// column group (GAS), and a CTA that finds that out has already paid the latency of decoding eight parent
// lists.  Per column group, keep the strings that have at least one parent in a block of the group.
// xoff: [string class][group] offset of the block in the operand, -1 if it does not exist.
void fill_group_klists(const f2_sigma& b, int spin, SSPlan& p, const std::vector<int64_t>& xoff, int ngroups) {
    const CIStringsHost& h = *b.S->host;
    p.gk_ptr.release();
    p.gk_list.release();
    p.gk_max = 0;
    if (h.blocks.size() <= 1 || ngroups <= 0)
        return;
    const DeviceLists& L = h.device_lists();
    const auto& cls = spin == 0 ? L.a_cls : L.b_cls;
    const auto& klist = L.h2_color_list[spin];
    const auto& ptr = L.h2_ptr[spin];
    const auto& ent = L.h2_ent[spin];
    const size_t ncls = xoff.size() / static_cast<size_t>(ngroups);
    std::vector<std::vector<int>> groups_of(ncls);
    for (size_t c = 0; c < ncls; ++c)
        for (int g = 0; g < ngroups; ++g)
            if (xoff[c * ngroups + g] >= 0)
                groups_of[c].push_back(g);
    std::vector<std::vector<uint32_t>> lists(ngroups);
    std::vector<int64_t> cls_seen(ncls, -1), grp_seen(ngroups, -1);
    for (size_t i = 0; i < klist.size(); ++i) {
        const uint32_t K = klist[i];
        for (uint32_t e = ptr[K]; e < ptr[K + 1]; ++e) {
            const uint32_t c = cls[ent[e] & 0xffffffffu];
            if (cls_seen[c] == static_cast<int64_t>(i))
                continue;
            cls_seen[c] = static_cast<int64_t>(i);
            for (int g : groups_of[c])
                if (grp_seen[g] != static_cast<int64_t>(i)) {
                    grp_seen[g] = static_cast<int64_t>(i);
                    lists[g].push_back(K);
                }
        }
    }
    std::vector<uint32_t> gptr(ngroups + 1, 0), flat;
    size_t total = 0;
    for (int g = 0; g < ngroups; ++g)
        total += lists[g].size();
    if (total >= (size_t(1) << 32))
        return; // keep the common list
    flat.reserve(total);
    for (int g = 0; g < ngroups; ++g) {
        gptr[g] = static_cast<uint32_t>(flat.size());
        flat.insert(flat.end(), lists[g].begin(), lists[g].end());
        p.gk_max = std::max<uint32_t>(p.gk_max, static_cast<uint32_t>(lists[g].size()));
    }
    gptr[ngroups] = static_cast<uint32_t>(flat.size());
    if (flat.empty())
        flat.push_back(0);
    p.gk_ptr.upload(gptr);
    p.gk_list.upload(flat);
}

// ownership: rank r of world owns rows [r nA / world, (r+1) nA / world) of every alpha class and,
// for the alpha same-spin exchange, the same kind of slice of every beta class's columns
std::unique_ptr<Part> make_part(f2_sigma& b, int rank, int world) {
    const CIStringsHost& h = *b.S->host;
    const DeviceLists& L = h.device_lists();
    const auto& A = h.space(0, K_STR_N);
    const auto& B = h.space(1, K_STR_N);
    const size_t nclsA = A.nclasses(), nclsB = B.nclasses();
    auto P = std::make_unique<Part>();
    P->rank = rank;
    P->world = world;
    P->lo.resize(nclsA);
    P->hi.resize(nclsA);
    P->cl.resize(nclsB);
    P->ch.resize(nclsB);
    for (size_t c = 0; c < nclsA; ++c) {
        const int64_t n = static_cast<int64_t>(A.strpcls(c));
        P->lo[c] = n * rank / world;
        P->hi[c] = n * (rank + 1) / world;
    }
    for (size_t c = 0; c < nclsB; ++c) {
        const int64_t n = static_cast<int64_t>(B.strpcls(c));
        P->cl[c] = n * rank / world;
        P->ch[c] = n * (rank + 1) / world;
    }
    std::vector<int64_t> sblk(nclsA * nclsB, -1);
    std::vector<uint32_t> srow0(nclsA + 1, 0);
    std::vector<CopyPiece> pieces;
    uint32_t cta0 = 0;
    for (size_t c = 0; c < nclsA; ++c)
        srow0[c] = static_cast<uint32_t>(P->lo[c]);
    for (const auto& blk : h.blocks) {
        const int64_t rows = P->hi[blk.cA] - P->lo[blk.cA];
        const int64_t cols = P->ch[blk.cB] - P->cl[blk.cB];
        P->loc_off.push_back(P->local_size);
        P->y_off.push_back(P->y_size);
        if (blk.nA * blk.nB > 0) {
            sblk[blk.cA * nclsB + blk.cB] = P->local_size;
            if (rows > 0) {
                CopyPiece pc;
                pc.src_off = blk.offset + P->lo[blk.cA] * blk.nB;
                pc.dst_off = P->local_size;
                pc.n = rows * blk.nB;
                pc.cta0 = cta0;
                cta0 += static_cast<uint32_t>((pc.n + 1023) / 1024);
                pieces.push_back(pc);
            }
            P->local_size += rows * blk.nB;
            P->y_size += blk.nA * cols;
        }
    }
    P->sblk_off.upload(sblk);
    P->srow0.upload(srow0);
    P->npieces = static_cast<int>(pieces.size());
    P->piece_ctas = cta0;
    P->pieces.upload(pieces);

    // ---- opposite-spin work items: (owned alpha string, beta class of the intermediate, m-split)
    if (b.t3_ks > 0) {
        std::vector<uint32_t> items;
        std::vector<int> kcount(nclsB);
        std::vector<char> row_has_block(nclsA, 0);
        for (const auto& blk : h.blocks)
            if (blk.nA * blk.nB > 0)
                row_has_block[blk.cA] = 1;
        for (size_t cA = 0; cA < nclsA; ++cA) {
            if (!row_has_block[cA])
                continue;
            for (int64_t i = P->lo[cA]; i < P->hi[cA]; ++i) {
                const uint32_t g = A.cls_off[cA] + static_cast<uint32_t>(i);
                std::fill(kcount.begin(), kcount.end(), 0);
                for (uint32_t e = L.as_ptr[g]; e < L.as_ptr[g + 1]; ++e) {
                    const uint32_t cJ = L.a_cls[L.as_ent[e] & 0xffffffffu];
                    for (size_t cB = 0; cB < nclsB; ++cB) {
                        const int64_t bi = h.block_of[cJ * nclsB + cB];
                        if (bi >= 0 && h.blocks[bi].nA * h.blocks[bi].nB > 0)
                            kcount[cB] += 1;
                    }
                }
                for (size_t cB = 0; cB < nclsB; ++cB) {
                    if (B.strpcls(cB) == 0 || kcount[cB] == 0)
                        continue;
                    P->t3_kcols += static_cast<int64_t>(kcount[cB]) * static_cast<int64_t>(B.strpcls(cB));
                    for (int ms = 0; ms < b.t3_msplit; ++ms) {
                        items.push_back(g);
                        items.push_back(static_cast<uint32_t>(cB) | (static_cast<uint32_t>(ms) << 16));
                    }
                }
            }
        }
        P->nitems = static_cast<int>(items.size() / 2);
        P->items.upload(items);
    }

    // ---- alpha same-spin plan: every row, owned columns; Y = sigma itself on one GPU, the exchange
    // buffer ([nA][owned columns] per block) otherwise
    {
        std::vector<int64_t> xoff(nclsA * nclsB, -1), yoff(nclsA * nclsB, -1);
        std::vector<uint32_t> ncols(nclsB), ldx(nclsB), ldy(nclsB), tiles;
        bool even = b.D->al16;
        for (size_t bi = 0; bi < h.blocks.size(); ++bi) {
            const auto& blk = h.blocks[bi];
            if (blk.nA * blk.nB == 0)
                continue;
            xoff[blk.cA * nclsB + blk.cB] = blk.offset + P->cl[blk.cB];
            yoff[blk.cA * nclsB + blk.cB] = world == 1 ? blk.offset : P->y_off[bi];
            even = even && (P->cl[blk.cB] % 2 == 0);
        }
        for (size_t c = 0; c < nclsB; ++c) {
            ncols[c] = static_cast<uint32_t>(P->ch[c] - P->cl[c]);
            ldx[c] = static_cast<uint32_t>(B.strpcls(c));
            ldy[c] = world == 1 ? ldx[c] : ncols[c];
        }
        fill_tiles(tiles, ncols);
        SSPlan& p = P->alpha_plan;
        p.ngroups = static_cast<int>(nclsB);
        p.ntiles = static_cast<int>(tiles.size() / 2);
        p.xoff.upload(xoff);
        p.yoff.upload(yoff);
        p.g_ncols.upload(ncols);
        p.g_ldx.upload(ldx);
        p.g_ldy.upload(ldy);
        p.tiles.upload(tiles);
        p.al16 = even;
        fill_group_klists(b, 0, p, xoff, static_cast<int>(nclsB));
    }
    // ---- rows form: the alpha same-spin plan whose operand is the exchanged column slab (laid out like the
    // exchange buffer: per block [nIa][owned columns]), and the reduce-scatter layout of the gather-form
    // opposite-spin kernels: world slots of slot_elems doubles, slot q = the local vector of rank q
    if (world > 1 && h.ndet > 0) {
        std::vector<int64_t> xoff(nclsA * nclsB, -1);
        std::vector<uint32_t> ncols(nclsB), tiles;
        for (size_t bi = 0; bi < h.blocks.size(); ++bi) {
            const auto& blk = h.blocks[bi];
            if (blk.nA * blk.nB > 0)
                xoff[blk.cA * nclsB + blk.cB] = P->y_off[bi];
        }
        for (size_t c = 0; c < nclsB; ++c)
            ncols[c] = static_cast<uint32_t>(P->ch[c] - P->cl[c]);
        fill_tiles(tiles, ncols);
        SSPlan& p = P->alpha_cols_plan;
        p.ngroups = static_cast<int>(nclsB);
        p.ntiles = static_cast<int>(tiles.size() / 2);
        p.xoff.upload(xoff);
        p.yoff.upload(xoff);
        p.g_ncols.upload(ncols);
        p.g_ldx.upload(ncols);
        p.g_ldy.upload(ncols);
        p.tiles.upload(tiles);
        p.al16 = false;
        fill_group_klists(b, 0, p, xoff, static_cast<int>(nclsB));
        // local layouts of every rank
        std::vector<std::vector<int64_t>> loc_off(world, std::vector<int64_t>(h.blocks.size(), 0));
        int64_t slot = 0;
        for (int q = 0; q < world; ++q) {
            int64_t pos = 0;
            for (size_t bi = 0; bi < h.blocks.size(); ++bi) {
                const auto& blk = h.blocks[bi];
                loc_off[q][bi] = pos;
                if (blk.nA * blk.nB > 0)
                    pos += (blk.nA * (q + 1) / world - blk.nA * q / world) * blk.nB;
            }
            slot = std::max(slot, pos);
        }
        P->slot_elems = slot;
        if (b.t3g && nclsA == 1 && nclsB == 1 && h.blocks.size() == 1) {
            const int64_t nA = static_cast<int64_t>(A.strpcls(0));
            int64_t rows = 0;
            for (int q = 0; q < world; ++q)
                rows = std::max(rows, nA * (q + 1) / world - nA * q / world);
            P->slot_rows = rows;
            std::vector<uint32_t> rowpad(static_cast<size_t>(nA));
            for (int q = 0; q < world; ++q)
                for (int64_t i = nA * q / world; i < nA * (q + 1) / world; ++i)
                    rowpad[static_cast<size_t>(i)] = static_cast<uint32_t>(q * rows + (i - nA * q / world));
            P->rowpad.upload(rowpad);
        } else if (b.t3gm && static_cast<int64_t>(world) * slot < (int64_t(1) << 32) - 1) {
            std::vector<uint32_t> jl, crow(static_cast<size_t>(h.nas) * nclsB, kNoRow),
                srow(static_cast<size_t>(h.nas) * nclsB, kNoRow);
            for (size_t cA = 0; cA < nclsA; ++cA) {
                const int64_t nA = static_cast<int64_t>(A.strpcls(cA));
                for (int64_t i = P->lo[cA]; i < P->hi[cA]; ++i)
                    jl.push_back(A.cls_off[cA] + static_cast<uint32_t>(i));
                for (size_t cB = 0; cB < nclsB; ++cB) {
                    const int64_t bi = h.block_of[cA * nclsB + cB];
                    if (bi < 0 || h.blocks[bi].nA * h.blocks[bi].nB == 0)
                        continue;
                    const int64_t nB = h.blocks[bi].nB;
                    for (int q = 0; q < world; ++q) {
                        const int64_t lo = nA * q / world, hi = nA * (q + 1) / world;
                        for (int64_t i = lo; i < hi; ++i) {
                            const size_t at = (A.cls_off[cA] + i) * nclsB + cB;
                            srow[at] = static_cast<uint32_t>(q * slot + loc_off[q][bi] + (i - lo) * nB);
                            if (q == rank)
                                crow[at] = static_cast<uint32_t>(loc_off[q][bi] + (i - lo) * nB);
                        }
                    }
                }
            }
            P->njlist = static_cast<int>(jl.size());
            P->jlist.upload(jl);
            P->crow_local.upload(crow);
            P->srow_slot.upload(srow);
        }
    }
    return P;
}

// beta same-spin work runs on transposed row slabs: a pass holds, for a set of (alpha class,
// row range) column groups, the transposed blocks [nB][rows] of every beta class paired with it.
// Both scratch buffers together must fit the set_memory budget (ci_sigma_builder.cc:70-78,
// ci_sigma_builder_knowles_handy.cc:15-32).
void plan_beta(f2_sigma& b, Part& P, int64_t cap_elems = 0) {
    const CIStringsHost& h = *b.S->host;
    P.beta_passes.clear();
    if (h.nb < 1 || h.ndet == 0) {
        P.beta_planned = true;
        return;
    }
    const size_t nclsA = h.nclsA(), nclsB = h.nclsB();
    int64_t cap = static_cast<int64_t>(b.memory_size / (2 * sizeof(double)));
    // columns (beta strings) attached to each alpha class
    std::vector<int64_t> width(nclsA, 0);
    for (const auto& blk : h.blocks)
        if (blk.nA * blk.nB > 0)
            width[blk.cA] += blk.nB;
    if (cap_elems > 0) { // a smaller slab on request, never below one row of the widest class
        int64_t wmax = 1;
        for (int64_t w : width)
            wmax = std::max(wmax, w);
        cap = std::min(cap, std::max(cap_elems, wmax));
    }
    struct Group {
        int cA;
        int64_t i0, rows;
    };
    std::vector<std::vector<Group>> passes;
    std::vector<Group> cur;
    int64_t used = 0;
    for (size_t cA = 0; cA < nclsA; ++cA) {
        if (width[cA] == 0 || P.hi[cA] == P.lo[cA])
            continue;
        if (cap < width[cA])
            throw std::runtime_error("The CI builder memory limit (" + std::to_string(b.memory_size) +
                                     " bytes) is too small for one K-block column of " +
                                     std::to_string(width[cA]) + " elements per buffer.");
        int64_t i0 = P.lo[cA];
        while (i0 < P.hi[cA]) {
            int64_t fit = (cap - used) / width[cA];
            if (fit == 0) {
                passes.push_back(cur);
                cur.clear();
                used = 0;
                continue;
            }
            const int64_t rows = std::min(fit, P.hi[cA] - i0);
            cur.push_back({static_cast<int>(cA), i0, rows});
            used += rows * width[cA];
            i0 += rows;
        }
    }
    if (!cur.empty())
        passes.push_back(cur);

    int64_t max_elems = 0;
    for (const auto& groups : passes) {
        auto bp = std::make_unique<BetaPass>();
        const int ng = static_cast<int>(groups.size());
        std::vector<int64_t> xoff(nclsB * ng, -1);
        std::vector<uint32_t> ncols(ng), tiles;
        std::vector<TrBlock> tr;
        int64_t pos = 0;
        uint32_t tile0 = 0;
        bool even = true;
        for (int g = 0; g < ng; ++g) {
            const Group& G = groups[g];
            ncols[g] = static_cast<uint32_t>(G.rows);
            for (size_t cB = 0; cB < nclsB; ++cB) {
                const int64_t bi = h.block_of[G.cA * nclsB + cB];
                if (bi < 0 || h.blocks[bi].nA * h.blocks[bi].nB == 0)
                    continue;
                const auto& blk = h.blocks[bi];
                xoff[cB * ng + g] = pos;
                TrBlock t;
                t.src_off = blk.offset + G.i0 * blk.nB;
                t.sig_off = P.world == 1 ? t.src_off : P.loc_off[bi] + (G.i0 - P.lo[G.cA]) * blk.nB;
                t.dst_off = pos;
                t.nA = static_cast<int32_t>(G.rows);
                t.nB = static_cast<int32_t>(blk.nB);
                t.ldA = blk.nB;
                t.tile0 = tile0;
                t.tiles_j = static_cast<uint32_t>((blk.nB + 31) / 32);
                tile0 += static_cast<uint32_t>((G.rows + 31) / 32) * t.tiles_j;
                tr.push_back(t);
                even = even && (pos % 2 == 0) && (G.rows % 2 == 0);
                pos += G.rows * blk.nB;
            }
        }
        fill_tiles(tiles, ncols);
        bp->plan.ngroups = ng;
        bp->plan.ntiles = static_cast<int>(tiles.size() / 2);
        bp->plan.xoff.upload(xoff);
        bp->plan.yoff.upload(xoff);
        bp->plan.g_ncols.upload(ncols);
        bp->plan.g_ldx.upload(ncols);
        bp->plan.g_ldy.upload(ncols);
        bp->plan.tiles.upload(tiles);
        bp->plan.al16 = even;
        fill_group_klists(b, 1, bp->plan, xoff, ng);
        bp->ntr = static_cast<int>(tr.size());
        bp->ntr_tiles = tile0;
        bp->tr.upload(tr);
        bp->tr_host = tr;
        bp->elems = pos;
        max_elems = std::max(max_elems, pos);
        P.beta_passes.push_back(std::move(bp));
    }
    if (static_cast<size_t>(std::max<int64_t>(max_elems, 1)) > b.XT.n) {
        b.XT.alloc(static_cast<size_t>(std::max<int64_t>(max_elems, 1)));
        b.YT.alloc(static_cast<size_t>(std::max<int64_t>(max_elems, 1)));
    }
    F2_CUDA(cudaDeviceSynchronize()); // the plan uploads (pageable memory) have landed before any stream reads them
    P.beta_planned = true;
}

void plan_hdiag(f2_sigma& b) {
    const CIStringsHost& h = *b.S->host;
    std::vector<HdiagBlock> hb;
    uint32_t tile0 = 0;
    for (const auto& blk : h.blocks) {
        if (blk.nA * blk.nB == 0)
            continue;
        HdiagBlock x;
        x.off = blk.offset;
        x.a0 = h.space(0, K_STR_N).cls_off[blk.cA];
        x.b0 = h.space(1, K_STR_N).cls_off[blk.cB];
        x.nA = static_cast<int32_t>(blk.nA);
        x.nB = static_cast<int32_t>(blk.nB);
        x.tile0 = tile0;
        tile0 += static_cast<uint32_t>(blk.nA);
        hb.push_back(x);
    }
    b.hd_nblocks = static_cast<int>(hb.size());
    b.hd_rows = tile0;
    b.hd_blocks.upload(hb);
}

template <int KS, bool AL16, int NS, bool TAIL, bool ONEB>
void launch_t3_oneb(const f2_sigma& b, const T3Params& p, int nitems, cudaStream_t st) {
    const dim3 grid(nitems), block(b.t3_warps * 32);
    F2_CUDA(cudaFuncSetAttribute(t3_kernel<KS, AL16, NS, TAIL, ONEB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(b.t3_smem)));
    t3_kernel<KS, AL16, NS, TAIL, ONEB><<<grid, block, b.t3_smem, st>>>(p);
    b.launches += 1;
}

template <int KS, bool AL16, int NS, bool TAIL>
void launch_t3_inst(const f2_sigma& b, const T3Params& p, int nitems, cudaStream_t st) {
    if constexpr (AL16) { // the single-beta-class fast emit needs the paired table loads
        if (p.nclsB == 1 && p.tgt_pair) {
            launch_t3_oneb<KS, AL16, NS, TAIL, true>(b, p, nitems, st);
            return;
        }
    }
    launch_t3_oneb<KS, AL16, NS, TAIL, false>(b, p, nitems, st);
}

template <int KS, bool AL16>
void launch_t3_al(const f2_sigma& b, const T3Params& p, int nitems, cudaStream_t st) {
    if (b.t3_tail) {
        if constexpr (KS <= 24) { // the tail strip costs KS more registers; 8 warps keep that affordable
            if (b.t3_ns == 3)
                launch_t3_inst<KS, AL16, 3, true>(b, p, nitems, st);
            else
                launch_t3_inst<KS, AL16, 2, true>(b, p, nitems, st);
            return;
        }
    }
    if (b.t3_ns == 3)
        launch_t3_inst<KS, AL16, 3, false>(b, p, nitems, st);
    else
        launch_t3_inst<KS, AL16, 2, false>(b, p, nitems, st);
}

template <int KS>
void launch_t3_ks(const f2_sigma& b, const T3Params& p, int nitems, cudaStream_t st) {
    if (b.D->al16 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0)
        launch_t3_al<KS, true>(b, p, nitems, st);
    else
        launch_t3_al<KS, false>(b, p, nitems, st);
}


template <int NM16, bool HAS8, int NX = 0>
void launch_t3g_inst(const f2_sigma& b, const T3GParams& p, int nrows, cudaStream_t st) {
    // 6 row tiles: 8 target columns per warp item instead of 16, which halves the accumulators and keeps 12 warps
    // per SM; F2_T3G_WIDE = the 16-column items on 8 warps (the round-1 shape)
    constexpr int THREADS = 384;
    const bool wide = std::getenv("F2_T3G_WIDE") != nullptr;
    if constexpr (NM16 >= 6) {
        if (wide) {
            if (b.t3g_long) {
                F2_CUDA(cudaFuncSetAttribute(t3g_kernel<NM16, HAS8, true, 256, 2>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(b.t3g_smem)));
                t3g_kernel<NM16, HAS8, true, 256, 2><<<nrows, 256, b.t3g_smem, st>>>(p);
            } else {
                F2_CUDA(cudaFuncSetAttribute(t3g_kernel<NM16, HAS8, false, 256, 2>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(b.t3g_smem)));
                t3g_kernel<NM16, HAS8, false, 256, 2><<<nrows, 256, b.t3g_smem, st>>>(p);
            }
            b.launches += 1;
            return;
        }
    }
    constexpr int NT = NM16 >= 6 ? 1 : 2;
    if (b.t3g_long) {
        F2_CUDA(cudaFuncSetAttribute(t3g_kernel<NM16, HAS8, true, THREADS, NT, NX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(b.t3g_smem)));
        t3g_kernel<NM16, HAS8, true, THREADS, NT, NX><<<nrows, THREADS, b.t3g_smem, st>>>(p);
    } else {
        F2_CUDA(cudaFuncSetAttribute(t3g_kernel<NM16, HAS8, false, THREADS, NT, NX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(b.t3g_smem)));
        t3g_kernel<NM16, HAS8, false, THREADS, NT, NX><<<nrows, THREADS, b.t3g_smem, st>>>(p);
    }
    b.launches += 1;
}

// gather-form opposite-spin term over the input alpha strings [first, first + nrows) (single block).
// C points at row 0 of the block (rows outside the range are never read); rowpad (may be NULL) maps
// output rows into the reduce-scatter layout of a sharded build.
void launch_t3g(const f2_sigma& b, const double* C, double* sigma, cudaStream_t st, int first = 0,
                int nrows = -1, const uint32_t* rowpad = nullptr) {
    const CIStringsHost& h = *b.S->host;
    const DevStrings& D = *b.D;
    T3GParams p;
    p.C = C;
    p.sigma = sigma;
    p.as_ptr = D.as_ptr.p;
    p.as_ent = D.as_ent.p;
    p.gsrc = b.gsrc.p;
    p.VxT = b.VxT.p;
    p.nbs = h.nbs;
    p.gld = b.t3g_gld;
    p.M = b.M;
    p.Mpad = b.t3g_mpad;
    p.norb = h.norb;
    p.first = first;
    p.astride = b.t3g_astride;
    p.rowpad = rowpad;
    p.nsm = b.t3g_nsm;
    if (nrows < 0)
        nrows = static_cast<int>(h.nas);
    if (nrows == 0)
        return;
    const int key = b.t3g_nm16 * 2 + b.t3g_has8 + 100 * b.t3g_nx;
    switch (key) {
    case 108: launch_t3g_inst<4, false, 1>(b, p, nrows, st); break;
    case 208: launch_t3g_inst<4, false, 2>(b, p, nrows, st); break;
    case 110: launch_t3g_inst<5, false, 1>(b, p, nrows, st); break;
    case 210: launch_t3g_inst<5, false, 2>(b, p, nrows, st); break;
    case 2: launch_t3g_inst<1, false>(b, p, nrows, st); break;
    case 3: launch_t3g_inst<1, true>(b, p, nrows, st); break;
    case 4: launch_t3g_inst<2, false>(b, p, nrows, st); break;
    case 5: launch_t3g_inst<2, true>(b, p, nrows, st); break;
    case 6: launch_t3g_inst<3, false>(b, p, nrows, st); break;
    case 7: launch_t3g_inst<3, true>(b, p, nrows, st); break;
    case 8: launch_t3g_inst<4, false>(b, p, nrows, st); break;
    case 9: launch_t3g_inst<4, true>(b, p, nrows, st); break;
    case 10: launch_t3g_inst<5, false>(b, p, nrows, st); break;
    case 11: launch_t3g_inst<5, true>(b, p, nrows, st); break;
    case 12: launch_t3g_inst<6, false>(b, p, nrows, st); break;
    default: throw std::runtime_error("CISigmaBuilder: unsupported gather-form shape");
    }
    F2_CUDA(cudaGetLastError());
}

// gather-form opposite-spin term for spaces with several determinant blocks.  One GPU: every alpha string,
// C and sigma in the CI-vector layout.  Sharded (P given): the owned input strings, C = the local vector,
// sigma = the reduce-scatter buffer in the slot layout.
template <int KB, int TG>
void launch_t3gp_inst(const T3GMParams& p, int nrows, size_t smem, cudaStream_t st) {
    constexpr int THREADS = T3GP_THREADS;
    F2_CUDA(cudaFuncSetAttribute(t3gp_kernel<THREADS, KB, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    t3gp_kernel<THREADS, KB, TG><<<nrows, THREADS, smem, st>>>(p);
}
void launch_t3gp_shape(int shape, const T3GMParams& p, int nrows, size_t smem, cudaStream_t st) {
    switch (shape) {
    case 1: launch_t3gp_inst<3, 2>(p, nrows, smem, st); break;
    case 2: launch_t3gp_inst<4, 1>(p, nrows, smem, st); break;
    case 3: launch_t3gp_inst<3, 1>(p, nrows, smem, st); break;
    case 4: launch_t3gp_inst<5, 1>(p, nrows, smem, st); break;
    case 5: launch_t3gp_inst<3, 3>(p, nrows, smem, st); break;
    default: launch_t3gp_inst<T3GP_KB, T3GP_TG>(p, nrows, smem, st); break;
    }
}

void launch_t3gm(const f2_sigma& b, const double* C, double* sigma, cudaStream_t st, const Part* P = nullptr) {
    const CIStringsHost& h = *b.S->host;
    const DevStrings& D = *b.D;
    T3GMParams p;
    p.C = C;
    p.sigma = sigma;
    p.as_ptr = D.as_ptr.p;
    p.as_ent = D.as_ent.p;
    p.a_cls = D.a_cls.p;
    p.gsrc = b.gm_gsrc.p;
    p.gpos = b.gm_gpos.p;
    p.nB_cls = D.nB_cls.p;
    p.crow = P ? P->crow_local.p : b.gm_crow.p;
    p.srow = P ? P->srow_slot.p : b.gm_crow.p;
    p.jlist = P ? P->jlist.p : nullptr;
    p.ml_ent = b.gm_ml_ent.p;
    p.ml_cnt = b.gm_ml_cnt.p;
    p.VxT = b.VxT.p;
    p.gld = b.gm_gld;
    p.M = b.M;
    p.Mpad = b.gm_mpad;
    p.norb = h.norb;
    p.nclsB = static_cast<int>(h.nclsB());
    p.first = 0;
    p.astride = b.gm_astride;
    p.rows = b.gm_nt16 * 16;
    const int nrows = P ? P->njlist : static_cast<int>(h.nas);
    if (nrows == 0)
        return;
    // shapes (KB, TG): (4, 2), (3, 2), (4, 1) take part in the timing; (3, 1), (5, 1), (3, 3) exist (no register spills at
    // 512 threads) but won on none of the measured spaces (profiles/r3f_spaces_t3gp_six_shapes.log) and are only reachable
    // through F2_T3GP_SHAPE=<0..5>, which pins a shape (tests, tuning).  Every shape computes the same sigma, so the first
    // builds of a builder simply take turns (launch 0 untimed, then one timed launch each) and the fastest one is kept.
    f2_sigma::T3gpTune& T = b.t3gp_tune;
    if (T.choice < 0) {
        if (const char* e = std::getenv("F2_T3GP_SHAPE")) {
            T.choice = std::min(std::max(std::atoi(e), 0), f2_sigma::T3gpTune::kShapes - 1);
        } else {
            if (T.pending) {
                F2_CUDA(cudaEventSynchronize(T.ev[1]));
                F2_CUDA(cudaEventElapsedTime(&T.ms[T.trial], T.ev[0], T.ev[1]));
                T.pending = false;
            }
            if (T.calls > f2_sigma::T3gpTune::kCand) {
                T.choice = static_cast<int>(std::min_element(T.ms, T.ms + f2_sigma::T3gpTune::kCand) - T.ms);
                if (std::getenv("F2_T3GP_TRACE"))
                    std::fprintf(stderr, "t3gp_kernel shapes (KB, TG) = (4, 2) %.3f, (3, 2) %.3f, (4, 1) %.3f ms: shape %d kept\n", T.ms[0],
                                 T.ms[1], T.ms[2], T.choice);
            }
        }
    }
    const bool timing = T.choice < 0 && T.calls >= 1;
    const int shape = T.choice >= 0 ? T.choice : (T.calls == 0 ? 0 : T.calls - 1);
    if (timing) {
        for (cudaEvent_t& e : T.ev)
            if (!e)
                F2_CUDA(cudaEventCreate(&e));
        F2_CUDA(cudaEventRecord(T.ev[0], st));
    }
    launch_t3gp_shape(shape, p, nrows, b.gm_smem, st);
    if (timing) {
        F2_CUDA(cudaEventRecord(T.ev[1], st));
        T.trial = shape;
        T.pending = true;
    }
    T.calls += 1;
    b.launches += 1;
    F2_CUDA(cudaGetLastError());
}

void launch_t3(const f2_sigma& b, const Part& P, const double* C, double* sigma, bool fold,
               cudaStream_t st) {
    if (P.nitems == 0)
        return;
    const CIStringsHost& h = *b.S->host;
    const DevStrings& D = *b.D;
    T3Params p;
    p.C = C;
    p.sigma = sigma;
    p.items = P.items.p;
    p.a_cls = D.a_cls.p;
    p.a_loc = D.a_loc.p;
    p.as_ptr = D.as_ptr.p;
    p.as_ent = D.as_ent.p;
    p.tgt = b.tgt.p;
    p.tgt_pair = b.tgt_pair;
    p.fold = fold ? 1 : 0;
    p.kk2 = b.kk2.p;
    p.rowmap = b.rowmap.p;
    p.nbs = h.nbs;
    p.blk_off = D.blk_off.p;
    p.sblk_off = P.sblk_off.p;
    p.srow0 = P.srow0.p;
    p.nB_cls = D.nB_cls.p;
    p.b_off = D.b_off.p;
    p.nclsA = static_cast<int>(h.nclsA());
    p.nclsB = static_cast<int>(h.nclsB());
    p.norb = h.norb;
    p.Vx = b.Vx.p;
    p.M = b.M;
    if (b.t3_generic) {
        const size_t smem = gen_smem_bytes(b.t3_gen_rows, b.t3_gen_tw, p.nclsB);
        F2_CUDA(cudaFuncSetAttribute(t3_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        t3_generic_kernel<<<P.nitems, GEN_THREADS, smem, st>>>(p, b.t3_gen_rows, b.t3_gen_tw);
        b.launches += 1;
        F2_CUDA(cudaGetLastError());
        return;
    }
    switch (b.t3_ks) {
    case 4: launch_t3_ks<4>(b, p, P.nitems, st); break;
    case 8: launch_t3_ks<8>(b, p, P.nitems, st); break;
    case 12: launch_t3_ks<12>(b, p, P.nitems, st); break;
    case 16: launch_t3_ks<16>(b, p, P.nitems, st); break;
    case 20: launch_t3_ks<20>(b, p, P.nitems, st); break;
    case 24: launch_t3_ks<24>(b, p, P.nitems, st); break;
    case 28: launch_t3_ks<28>(b, p, P.nitems, st); break;
    case 32: launch_t3_ks<32>(b, p, P.nitems, st); break;
    default: throw std::runtime_error("CISigmaBuilder: unsupported opposite-spin K size");
    }
    F2_CUDA(cudaGetLastError());
}

// ssb_kernel (parent rows staged in shared memory, same_spin_staged.cuh) or ssg_kernel (parent rows through registers).
// Measured (profiles/r2x_quick.log): ssb_kernel is 24 % faster on symmetry-blocked spaces (CAS(16,16) D2h: 3.63 against
// 4.78 ms per spin), where strings without parents in a column group and short lists break ssg_kernel's register
// pipeline; on one determinant block both sit at the same L2 reduction rate and ssg_kernel is 4 % faster (19.4 / 20.2 ms
// at CAS(16,16)).  Default: ssb_kernel for spaces with per-group string lists and up to 56 parent rows per string (two CTAs of
// it per SM); CAS(16,16) C2v: 5.48 against 6.25 ms per spin (profiles/r3a_quick_ssb_16_warps.log).
// F2_NO_SSB: never; F2_SSB_ALWAYS: wherever it fits.
inline int staged_same_spin_mode() {
    return std::getenv("F2_NO_SSB") ? 0 : (std::getenv("F2_SSB_ALWAYS") ? 2 : 1); // read per launch: a test hook
}

template <int NM16, bool HAS8, bool AL16, bool GK>
void launch_ssg_variant(const SSParams& p, const SSGroupLists& gk, dim3 grid, cudaStream_t st) {
    constexpr int rows_b = NM16 * 16 + (HAS8 ? 8 : 0);
    if constexpr (ssb_fits(rows_b)) {
        if (staged_same_spin_mode() == 2 || (staged_same_spin_mode() == 1 && GK)) {
            constexpr size_t smem_b = ssb_smem_bytes(rows_b, ssb_warps(rows_b));
            if (smem_b > 48 * 1024)
                F2_CUDA(cudaFuncSetAttribute(ssb_kernel<NM16, HAS8, AL16, GK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem_b)));
            ssb_kernel<NM16, HAS8, AL16, GK><<<grid, 32 * ssb_warps(rows_b), smem_b, st>>>(p, gk);
            return;
        }
    }
    constexpr size_t smem = ssg_smem_bytes(NM16 * 16 + (HAS8 ? 8 : 0));
    if (smem > 48 * 1024)
        F2_CUDA(cudaFuncSetAttribute(ssg_kernel<NM16, HAS8, AL16, GK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
    ssg_kernel<NM16, HAS8, AL16, GK><<<grid, 256, smem, st>>>(p, gk);
}

// GK (per-group string lists) is a separate instance and a separate kernel argument: the common-list kernel
// keeps the register allocation it was tuned with (114 registers at CAS(16,16); 125 with two more fields in
// SSParams, measured 4 % slower)
template <int NM16, bool HAS8>
void launch_ssg_inst(const SSParams& p, const SSGroupLists& gk, bool al16, dim3 grid, cudaStream_t st) {
    if (gk.ptr) {
        if (al16)
            launch_ssg_variant<NM16, HAS8, true, true>(p, gk, grid, st);
        else
            launch_ssg_variant<NM16, HAS8, false, true>(p, gk, grid, st);
    } else {
        if (al16)
            launch_ssg_variant<NM16, HAS8, true, false>(p, gk, grid, st);
        else
            launch_ssg_variant<NM16, HAS8, false, false>(p, gk, grid, st);
    }
}

// every (N-2)-electron string of one spin for one plan, one launch (reductions into Y commute)
void launch_same_spin(const f2_sigma& b, int spin, const SSPlan& plan, const double* X, double* Y,
                      const double* W, cudaStream_t st) {
    const CIStringsHost& h = *b.S->host;
    const DeviceLists& L = h.device_lists();
    const DevStrings& D = *b.D;
    const size_t nk = L.h2_color_list[spin].size();
    if (plan.ntiles == 0 || nk == 0)
        return;
    SSParams p;
    p.X = X;
    p.Y = Y;
    p.h2_ptr = D.h2_ptr[spin].p;
    p.h2_ent = D.h2_ent[spin].p;
    p.s_cls = spin == 0 ? D.a_cls.p : D.b_cls.p;
    p.s_loc = spin == 0 ? D.a_loc.p : D.b_loc.p;
    p.xoff = plan.xoff.p;
    p.yoff = plan.yoff.p;
    p.g_ncols = plan.g_ncols.p;
    p.g_ldx = plan.g_ldx.p;
    p.g_ldy = plan.g_ldy.p;
    p.W = W;
    p.P = b.P[spin];
    p.ngroups = plan.ngroups;
    const bool al16 = plan.al16 && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) & 15) == 0;
    const int mmax = std::max(L.h2_max[spin], 1);
    const int nm16 = mmax / 16 + (mmax % 16 > 8 ? 1 : 0), has8 = (mmax % 16 > 0 && mmax % 16 <= 8) ? 1 : 0;
    const SSGroupLists gk{plan.gk_ptr.p, plan.gk_list.p};
    if (gk.ptr && plan.gk_max == 0)
        return; // no string has a parent in any block of this plan
    // grid: x = K (fastest: all K of a column panel run together, the panel stays in L2), y = panel
    const size_t nk_grid = gk.ptr ? plan.gk_max : nk;
    if (b.ss_generic[spin]) {
        const int rows = b.ss_gen_rows[spin], tw = b.ss_gen_tw[spin];
        const size_t smem = gen_smem_bytes(rows, tw);
        F2_CUDA(cudaFuncSetAttribute(ss_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        F2_CUDA(cudaFuncSetAttribute(ss_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        for (size_t k0 = 0; k0 < nk_grid; k0 += size_t(1) << 30) {
            const uint32_t nkk = static_cast<uint32_t>(std::min<size_t>(size_t(1) << 30, nk_grid - k0));
            p.klist = D.h2_clist[spin].p + k0;
            p.nk = static_cast<int>(nkk);
            SSGroupLists gk0 = gk;
            if (gk0.ptr && k0 > 0)
                throw std::runtime_error("CISigmaBuilder: more than 2^30 two-hole strings per column group");
            for (int t0 = 0; t0 < plan.ntiles; t0 += 65535) {
                const int nt = std::min(65535, plan.ntiles - t0);
                p.tiles = plan.tiles.p + 2 * static_cast<size_t>(t0);
                const dim3 grid(nkk, nt);
                b.launches += 1;
                if (gk0.ptr)
                    ss_generic_kernel<true><<<grid, GEN_THREADS, smem, st>>>(p, gk0, rows, tw);
                else
                    ss_generic_kernel<false><<<grid, GEN_THREADS, smem, st>>>(p, gk0, rows, tw);
            }
        }
        F2_CUDA(cudaGetLastError());
        return;
    }
    for (size_t k0 = 0; k0 < nk_grid; k0 += size_t(1) << 30) {
        const uint32_t nkk = static_cast<uint32_t>(std::min<size_t>(size_t(1) << 30, nk_grid - k0));
        p.klist = D.h2_clist[spin].p + k0;
        p.nk = static_cast<int>(nkk);
        for (int t0 = 0; t0 < plan.ntiles; t0 += 65535) {
            const int nt = std::min(65535, plan.ntiles - t0);
            p.tiles = plan.tiles.p + 2 * static_cast<size_t>(t0);
            const dim3 grid((nkk + SSG_KG - 1) / SSG_KG, nt);
            b.launches += 1;
            switch (nm16 * 2 + has8) {
            case 1: launch_ssg_inst<0, true>(p, gk, al16, grid, st); break;
            case 2: launch_ssg_inst<1, false>(p, gk, al16, grid, st); break;
            case 3: launch_ssg_inst<1, true>(p, gk, al16, grid, st); break;
            case 4: launch_ssg_inst<2, false>(p, gk, al16, grid, st); break;
            case 5: launch_ssg_inst<2, true>(p, gk, al16, grid, st); break;
            case 6: launch_ssg_inst<3, false>(p, gk, al16, grid, st); break;
            case 7: launch_ssg_inst<3, true>(p, gk, al16, grid, st); break;
            case 8: launch_ssg_inst<4, false>(p, gk, al16, grid, st); break;
            case 9: launch_ssg_inst<4, true>(p, gk, al16, grid, st); break;
            case 10: launch_ssg_inst<5, false>(p, gk, al16, grid, st); break;
            case 11: launch_ssg_inst<5, true>(p, gk, al16, grid, st); break;
            case 12: launch_ssg_inst<6, false>(p, gk, al16, grid, st); break;
            case 13: launch_ssg_inst<6, true>(p, gk, al16, grid, st); break;
            case 14: launch_ssg_inst<7, false>(p, gk, al16, grid, st); break;
            case 15: launch_ssg_inst<7, true>(p, gk, al16, grid, st); break;
            case 16: launch_ssg_inst<8, false>(p, gk, al16, grid, st); break;
            default: throw std::runtime_error("CISigmaBuilder: unsupported same-spin block size");
            }
        }
    }
    F2_CUDA(cudaGetLastError());
}

void collect_set(f2_sigma& b, int set) {
    if (!b.set_pending[set])
        return;
    cudaEvent_t* ev = b.evs[set];
    F2_CUDA(cudaEventSynchronize(ev[4]));
    float ms[4];
    for (int i = 0; i < 4; ++i)
        F2_CUDA(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    float tot;
    F2_CUDA(cudaEventElapsedTime(&tot, ev[0], ev[4]));
    if (b.t3_first[set]) { // intervals are [opposite spin, H0, alpha, beta]
        const float t3 = ms[0];
        ms[0] = ms[1];
        ms[1] = ms[2];
        ms[2] = ms[3];
        ms[3] = t3;
    }
    b.last_ms[0] = ms[0];
    b.last_ms[1] = ms[1];
    b.last_ms[2] = ms[2];
    b.last_ms[3] = ms[3];
    b.last_ms[4] = tot;
    b.t_acc[0] += tot * 1e-3;
    b.t_acc[1] += ms[3] * 1e-3;
    b.t_acc[2] += ms[1] * 1e-3;
    b.t_acc[3] += ms[2] * 1e-3;
    b.builds += 1;
    b.set_pending[set] = false;
}

// every build that has not been accounted for yet, oldest first (waits for the newest one)
void collect_times(f2_sigma& b) {
    for (int k = 1; k <= f2_sigma::kEvSets; ++k)
        collect_set(b, (b.ev_idx + k) % f2_sigma::kEvSets);
}

// next event set of the ring; the build that used it kEvSets builds ago is accounted for first
void begin_build(f2_sigma& b) {
    b.ev_idx = (b.ev_idx + 1) % f2_sigma::kEvSets;
    collect_set(b, b.ev_idx);
    b.ev = b.evs[b.ev_idx];
    b.t3_first[b.ev_idx] = false;
}

// sigma (the owned rows) from the full trial vector.  On one GPU yexch is unused; otherwise the
// alpha same-spin term of the owned columns of every row is left in yexch for the caller to exchange.
// t3_streamed (one GPU, host path): the gather-form opposite-spin term has already been added into a
// zeroed sigma while C was arriving; H0 then accumulates and the opposite-spin launch is skipped.
// after_beta_pass (host path): called after the launches of every beta row slab; with the streamed
// opposite-spin term that slab's sigma rows are final once those launches have run.
void apply(f2_sigma& b, Part& P, const double* C, double* sigma, double* yexch, cudaStream_t st,
           bool t3_streamed = false, const std::function<void(const BetaPass&)>* after_beta_pass = nullptr) {
    const CIStringsHost& h = *b.S->host;
    F2_CUDA(cudaSetDevice(b.device));
    begin_build(b);
    if (!P.beta_planned)
        plan_beta(b, P);
    if (h.ndet == 0)
        return;
    b.launches = 0;
    NvtxRange nvtx_build("f2_sigma build");
    NvtxPhases nvtx;
    nvtx.next("H0 (scalar energy)");
    F2_CUDA(cudaEventRecord(b.ev[0], st));
    if (P.world == 1) {
        if (t3_streamed)
            scale_copy_kernel<true><<<148 * 8, 256, 0, st>>>(C, sigma, b.E, h.ndet);
        else
            scale_copy_kernel<false><<<148 * 8, 256, 0, st>>>(C, sigma, b.E, h.ndet);
        b.launches += 1;
    } else if (P.piece_ctas) {
        piece_copy_kernel<false><<<P.piece_ctas, 256, 0, st>>>(C, sigma, P.pieces.p, P.npieces, b.E);
        b.launches += 1;
    }
    F2_CUDA(cudaEventRecord(b.ev[1], st));
    nvtx.next("same-spin alpha");
    const bool gather = b.t3g && P.world == 1 && P.nitems > 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
    const bool gather_m = !gather && b.t3gm && P.world == 1 && P.nitems > 0;
    const bool fold = !gather && !gather_m && b.fold_ok && P.world == 1 && P.nitems > 0;
    if (t3_streamed && !gather)
        throw std::logic_error("streamed opposite-spin term without the gather form");
    if (h.na >= 1 && !fold) {
        if (P.world > 1)
            F2_CUDA(cudaMemsetAsync(yexch, 0, static_cast<size_t>(P.y_size) * sizeof(double), st));
        launch_same_spin(b, 0, P.alpha_plan, C, P.world == 1 ? sigma : yexch, b.W[0].p, st);
    }
    F2_CUDA(cudaEventRecord(b.ev[2], st));
    nvtx.next("same-spin beta (transposed row slabs)");
    for (const auto& bp : P.beta_passes) {
        if (bp->ntr_tiles == 0)
            continue;
        transpose_kernel<false><<<bp->ntr_tiles, 256, 0, st>>>(C, b.XT.p, bp->tr.p, bp->ntr, false);
        b.launches += 2;
        F2_CUDA(cudaMemsetAsync(b.YT.p, 0, static_cast<size_t>(bp->elems) * sizeof(double), st));
        launch_same_spin(b, 1, bp->plan, b.XT.p, b.YT.p, fold ? b.Wfold.p : b.W[1].p, st);
        transpose_kernel<true><<<bp->ntr_tiles, 256, 0, st>>>(b.YT.p, sigma, bp->tr.p, bp->ntr, false);
        if (after_beta_pass)
            (*after_beta_pass)(*bp);
    }
    F2_CUDA(cudaEventRecord(b.ev[3], st));
    nvtx.next("opposite spin");
    if (t3_streamed)
        ; // done by the caller, piece by piece
    else if (gather)
        launch_t3g(b, C, sigma, st);
    else if (gather_m)
        launch_t3gm(b, C, sigma, st);
    else
        launch_t3(b, P, C, sigma, fold, st);
    F2_CUDA(cudaEventRecord(b.ev[4], st));
    F2_CUDA(cudaGetLastError());
    b.set_pending[b.ev_idx] = true;
}

void apply(f2_sigma& b, const double* C, double* sigma, int64_t n, cudaStream_t st) {
    const CIStringsHost& h = *b.S->host;
    if (n != h.ndet)
        throw std::invalid_argument("CISigmaBuilder.Hamiltonian: vectors must have ndet = " +
                                    std::to_string(h.ndet) + " elements, got " + std::to_string(n));
    apply(b, *b.single, C, sigma, nullptr, st);
}

// Rows form of the sharded build (gather-form opposite-spin kernels: t3g_kernel for one determinant block,
// t3gp_kernel for several): no rank ever holds the full trial vector.
//   sigma_loc  = E C_loc + beta same-spin term of the owned rows                 (complete)
//   ycols      = alpha same-spin term of [all rows][owned columns], from xcols    (caller: all-to-all, add)
//   scatter    = opposite-spin contributions of the owned INPUT rows to every sigma row, world slots of
//                slot_elems doubles, slot q laid out like the local vector of rank q
//                (caller: reduce-scatter, add)
// phases: 1 = the opposite-spin term into `scatter` (first, so that the caller can start the
// reduce-scatter while phase 2 runs), 2 = H0 and the two same-spin terms, 3 = both.
void apply_rows(f2_sigma& b, Part& P, const double* Cloc, const double* xcols, double* sigma, double* ycols,
                double* scatter, int phases, cudaStream_t st) {
    const CIStringsHost& h = *b.S->host;
    F2_CUDA(cudaSetDevice(b.device));
    if (phases & 1) {
        begin_build(b);
        b.t3_first[b.ev_idx] = true;
        b.launches = 0;
    }
    if (!P.beta_planned)
        plan_beta(b, P);
    if (h.ndet == 0)
        return;
    // event order of this form: start, [opposite spin], e1, [H0], e2, [alpha], e3, [beta], e4
    NvtxPhases nvtx;
    if (phases & 1) {
        nvtx.next("rows form: opposite spin into the scatter slots");
        F2_CUDA(cudaEventRecord(b.ev[0], st));
        F2_CUDA(cudaMemsetAsync(scatter, 0, static_cast<size_t>(P.world) * P.slot_elems * sizeof(double), st));
        if (h.na >= 1 && h.nb >= 1 && P.local_size > 0) {
            if (b.t3g) {
                const int64_t rows = P.hi[0] - P.lo[0];
                launch_t3g(b, Cloc - P.lo[0] * h.nbs, scatter, st, static_cast<int>(P.lo[0]), static_cast<int>(rows),
                           P.rowpad.p);
            } else {
                launch_t3gm(b, Cloc, scatter, st, &P);
            }
        }
        F2_CUDA(cudaEventRecord(b.ev[1], st));
    }
    // 2 = 4 | 8: H0 and both same-spin terms; 4 = H0 and the alpha same-spin term (needs xcols); 8 = the beta same-spin
    // term (needs only the local rows), after 4
    if (phases & (2 | 4)) {
        nvtx.next("rows form: H0 + same-spin alpha on the column slab");
        if (P.local_size > 0) {
            scale_copy_kernel<false><<<148 * 8, 256, 0, st>>>(Cloc, sigma, b.E, P.local_size);
            b.launches += 1;
        }
        F2_CUDA(cudaEventRecord(b.ev[2], st));
        F2_CUDA(cudaMemsetAsync(ycols, 0, static_cast<size_t>(std::max<int64_t>(P.y_size, 1)) * sizeof(double), st));
        if (h.na >= 1)
            launch_same_spin(b, 0, P.alpha_cols_plan, xcols, ycols, b.W[0].p, st);
        F2_CUDA(cudaEventRecord(b.ev[3], st));
    }
    if (phases & (2 | 8)) {
        nvtx.next("rows form: same-spin beta on the owned rows");
        for (const auto& bp : P.beta_passes) {
            if (bp->ntr_tiles == 0)
                continue;
            transpose_kernel<false><<<bp->ntr_tiles, 256, 0, st>>>(Cloc, b.XT.p, bp->tr.p, bp->ntr, true);
            b.launches += 2;
            F2_CUDA(cudaMemsetAsync(b.YT.p, 0, static_cast<size_t>(bp->elems) * sizeof(double), st));
            launch_same_spin(b, 1, bp->plan, b.XT.p, b.YT.p, b.W[1].p, st);
            transpose_kernel<true><<<bp->ntr_tiles, 256, 0, st>>>(b.YT.p, sigma, bp->tr.p, bp->ntr, false);
        }
        F2_CUDA(cudaEventRecord(b.ev[4], st));
        b.set_pending[b.ev_idx] = true;
    }
    F2_CUDA(cudaGetLastError());
}

// Phase 1 of apply_rows in pieces (one determinant block, t3g_kernel): the opposite-spin term of the local input rows
// [row0, row0 + nrows) only, so that a host-resident shard can be multiplied while its later rows are still crossing
// PCIe.  flags: 1 = first piece of a build (event set, scatter zeroed), 2 = last piece (closes the term's time).
void apply_rows_t3_piece(f2_sigma& b, Part& P, const double* Cloc, double* scatter, int64_t row0, int64_t nrows, int flags,
                         cudaStream_t st) {
    const CIStringsHost& h = *b.S->host;
    F2_CUDA(cudaSetDevice(b.device));
    if (!b.t3g)
        throw std::runtime_error("CISigmaBuilder: the opposite-spin term in pieces needs the one-block gather kernel");
    const int64_t rows = P.hi[0] - P.lo[0];
    if (row0 < 0 || nrows < 0 || row0 + nrows > rows)
        throw std::invalid_argument("CISigmaBuilder: piece outside the local rows");
    NvtxRange nvtx("rows form: opposite spin, one piece of the local rows");
    if (flags & 1) {
        begin_build(b);
        b.t3_first[b.ev_idx] = true;
        b.launches = 0;
        if (!P.beta_planned)
            plan_beta(b, P);
        F2_CUDA(cudaEventRecord(b.ev[0], st));
        F2_CUDA(cudaMemsetAsync(scatter, 0, static_cast<size_t>(P.world) * P.slot_elems * sizeof(double), st));
    }
    if (h.na >= 1 && h.nb >= 1 && nrows > 0)
        launch_t3g(b, Cloc - P.lo[0] * h.nbs, scatter, st, static_cast<int>(P.lo[0] + row0), static_cast<int>(nrows), P.rowpad.p);
    if (flags & 2)
        F2_CUDA(cudaEventRecord(b.ev[1], st));
    F2_CUDA(cudaGetLastError());
}

// the rows form needs the gather-form kernel and 16-byte aligned row slabs (even nbs)
bool rows_form_ok(const f2_sigma& b) {
    const CIStringsHost& h = *b.S->host;
    if (std::getenv("F2_NO_ROWS_FORM") != nullptr)
        return false;
    if (b.t3g)
        return h.blocks.size() == 1 && h.nclsA() == 1 && h.nclsB() == 1 && h.nbs % 2 == 0;
    return b.t3gm; // several blocks: the block-aware gather kernel (the slot layout must fit 32-bit offsets)
}

Part& shard_of(f2_sigma& b, int rank, int world) {
    if (world < 1 || rank < 0 || rank >= world)
        throw std::invalid_argument("CISigmaBuilder: invalid rank/world");
    if (world == 1)
        return *b.single;
    if (!b.shard || b.shard->rank != rank || b.shard->world != world) {
        F2_CUDA(cudaSetDevice(b.device));
        F2_CUDA(cudaDeviceSynchronize());
        b.shard = make_part(b, rank, world);
        F2_CUDA(cudaDeviceSynchronize()); // the uploads (pageable memory) have landed before any stream reads them
    }
    return *b.shard;
}

} // namespace

extern "C" {

f2_sigma* f2_sigma_create(f2_strings* strings, double E, const double* H, const double* V,
                          int log_level, int device) {
    try {
        int count = 0;
        cudaError_t err = cudaGetDeviceCount(&count);
        if (err != cudaSuccess || count == 0)
            throw CudaError("forte2_b200: no CUDA device available (the sigma builder has no CPU "
                            "fallback)");
        F2_CUDA(cudaSetDevice(device));
        auto b = std::make_unique<f2_sigma>();
        b->S = strings;
        b->device = device;
        b->log_level = log_level;
        b->norb = strings->host->norb;
        const bool trace = std::getenv("F2_TIME_LISTS") != nullptr;
        auto T0 = std::chrono::steady_clock::now();
        auto lap = [&](const char* what) {
            if (trace) {
                cudaDeviceSynchronize();
                const auto t = std::chrono::steady_clock::now();
                std::fprintf(stderr, "builder: %-28s %.3f s\n", what, std::chrono::duration<double>(t - T0).count());
                T0 = t;
            }
        };
        b->D = device_strings(strings, device);
        lap("lists on the device");
        for (auto& set : b->evs)
            for (auto& e : set)
                F2_CUDA(cudaEventCreate(&e));
        build_integrals(*b, E, H, V);
        lap("integrals + beta tables");
        const DeviceLists& L = strings->host->device_lists();
        for (int s = 0; s < 2; ++s) {
            // the tensor-core kernel keeps two m x m integral blocks in shared memory: up to 112 parents
            const int mmax = std::max(L.h2_max[s], 1);
            const int rows16 = (mmax / 16 + (mmax % 16 > 8 ? 1 : 0)) * 16 + ((mmax % 16 > 0 && mmax % 16 <= 8) ? 8 : 0);
            b->ss_generic[s] = rows16 > 128 || ssg_smem_bytes(rows16) > size_t(227) * 1024 - 1024 ||
                               std::getenv("F2_GENERIC_KERNELS") != nullptr;
            if (b->ss_generic[s]) {
                b->ss_gen_rows[s] = mmax;
                b->ss_gen_tw[s] = generic_strip(mmax, 0);
                if (b->ss_gen_tw[s] == 0)
                    throw std::runtime_error("CISigmaBuilder: " + std::to_string(mmax) + " parents per two-hole string "
                                             "do not fit the same-spin kernels");
            }
            b->ss_ks[s] = round_ks(std::min(mmax, 128));
        }
        plan_t3_shape(*b);
        lap("opposite-spin shape + row tables");
        b->single = make_part(*b, 0, 1);
        lap("partition, items, K lists");
        plan_hdiag(*b);
        lap("diagonal plan");
        F2_CUDA(cudaDeviceSynchronize()); // every table upload (pageable memory) has landed before any stream reads it
        return b.release();
    } catch (const std::exception& e) {
        status_of(e);
        return nullptr;
    }
}

void f2_sigma_destroy(f2_sigma* h) { delete h; }

int f2_sigma_set_hamiltonian(f2_sigma* h, double E, const double* H, const double* V) {
    try {
        F2_CUDA(cudaSetDevice(h->device));
        F2_CUDA(cudaDeviceSynchronize());
        const bool was = h->compressed;
        build_integrals(*h, E, H, V);
        if (was != h->compressed) {
            plan_t3_shape(*h);
            h->single = make_part(*h, 0, 1);
            h->shard.reset();
        }
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_set_algorithm(f2_sigma* h, const char* name) {
    const std::string a(name);
    if (a == "kh" || a == "knowles-handy")
        h->algorithm = "Knowles-Handy";
    else if (a == "hz" || a == "harrison-zarrabian")
        h->algorithm = "Harrison-Zarrabian";
    else {
        g_last_error = "CI algorithm " + a + " not valid.";
        return F2_ERUNTIME;
    }
    return F2_OK;
}

const char* f2_sigma_get_algorithm(const f2_sigma* h) { return h->algorithm.c_str(); }

int f2_strings_build_lists(f2_strings* S, int device) {
    try {
        if (!S)
            throw std::invalid_argument("null argument");
        std::lock_guard<std::mutex> lock(S->dev_mutex);
        const CIStringsHost& h = *S->host;
        if (h.lists_built())
            return F2_OK;
        if (device < 0) {
            h.device_lists();
            return F2_OK;
        }
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || device >= count)
            throw std::runtime_error("f2_strings_build_lists: no CUDA device " + std::to_string(device));
        F2_CUDA(cudaSetDevice(device));
        build_lists_on_device(h);
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_set_memory(f2_sigma* h, int64_t mb) {
    if (mb < 0) {
        g_last_error = "CI builder memory must be non-negative.";
        return F2_EINVAL;
    }
    h->memory_size = static_cast<size_t>(mb) * 1024 * 1024;
    // like the reference, the scratch is released here and re-planned lazily by the next build,
    // which is where a too-small budget is reported
    for (Part* P : {h->single.get(), h->shard.get(), h->host_part.get()})
        if (P) {
            P->beta_planned = false;
            P->beta_passes.clear();
        }
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    h->XT.release();
    h->YT.release();
    return F2_OK;
}

int f2_sigma_set_log_level(f2_sigma* h, int level) {
    h->log_level = level;
    return F2_OK;
}

int f2_sigma_apply(f2_sigma* h, const double* C_dev, double* sigma_dev, int64_t n, void* stream) {
    try {
        apply(*h, C_dev, sigma_dev, n, static_cast<cudaStream_t>(stream));
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_apply_batch(f2_sigma* h, const double* C_dev, double* sigma_dev, int64_t n, int nvec, int64_t ld,
                         void* stream) {
    try {
        if (nvec < 0 || (nvec > 1 && ld < n))
            throw std::invalid_argument("CISigmaBuilder.Hamiltonian: need nvec >= 0 and ld >= ndet");
        for (int v = 0; v < nvec; ++v)
            apply(*h, C_dev + static_cast<int64_t>(v) * ld, sigma_dev + static_cast<int64_t>(v) * ld, n,
                  static_cast<cudaStream_t>(stream));
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_apply_host(f2_sigma* h, const double* C_host, double* sigma_host, int64_t n) {
    try {
        NvtxRange nvtx("f2_sigma_apply_host (numpy to numpy)");
        const CIStringsHost& s = *h->S->host;
        if (n != s.ndet)
            throw std::invalid_argument("CISigmaBuilder.Hamiltonian: vectors must have ndet = " +
                                        std::to_string(s.ndet) + " elements, got " +
                                        std::to_string(n));
        F2_CUDA(cudaSetDevice(h->device));
        if (!h->own_stream)
            F2_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
        const bool trace = std::getenv("F2_TIME_HOST") != nullptr;
        const auto T0 = std::chrono::steady_clock::now();
        auto stamp = [&](const char* what) {
            if (trace)
                std::fprintf(stderr, "host path: %-34s %8.2f ms\n", what,
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count());
        };
        const size_t bytes = static_cast<size_t>(std::max<int64_t>(n, 1)) * sizeof(double);
        if (!h->hC) {
            F2_CUDA(cudaMallocHost(&h->hC, bytes));
            F2_CUDA(cudaMallocHost(&h->hS, bytes));
        }
        h->dC.alloc(static_cast<size_t>(std::max<int64_t>(n, 1)));
        h->dS.alloc(static_cast<size_t>(std::max<int64_t>(n, 1)));
        // pageable user memory -> pinned staging -> device in 64 MB pieces on a copy stream: the host copy
        // of piece i + 1 (several threads) overlaps the DMA of piece i.  With the gather-form kernel the
        // opposite-spin term of the alpha rows that have arrived is launched every few pieces, so the
        // tensor work starts while most of C is still crossing PCIe; the way back mirrors the staging.
        size_t piece = size_t(8) << 20;
        if (const char* e = std::getenv("F2_HOST_PIECE")) // test hook: piece size in doubles
            piece = std::max<size_t>(16, static_cast<size_t>(std::atoll(e)));
        const size_t total = static_cast<size_t>(n);
        const bool streamed = h->t3g && h->single->nitems > 0 && s.nbs > 0 &&
                              std::getenv("F2_NO_STREAMED_HOST") == nullptr;
        if (!h->copy_stream)
            F2_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        if (!h->copy_ev)
            F2_CUDA(cudaEventCreateWithFlags(&h->copy_ev, cudaEventDisableTiming));
        // the streamed opposite-spin launches alternate between two streams: the last, partly filled wave of CTAs of one
        // launch then runs beside the first waves of the next instead of in front of them
        if (!h->own_stream2)
            F2_CUDA(cudaStreamCreateWithFlags(&h->own_stream2, cudaStreamNonBlocking));
        if (!h->t3_ev)
            F2_CUDA(cudaEventCreateWithFlags(&h->t3_ev, cudaEventDisableTiming));
        if (streamed) {
            F2_CUDA(cudaMemsetAsync(h->dS.p, 0, bytes, h->own_stream));
            F2_CUDA(cudaEventRecord(h->t3_ev, h->own_stream));
            F2_CUDA(cudaStreamWaitEvent(h->own_stream2, h->t3_ev, 0));
        }
        int64_t rows_done = 0;
        int64_t streamed_launches = 0;
        // launch granularity: the first opposite-spin launch follows the first piece (the GPU starts after 64 MB, not
        // 256), later ones grow to four pieces (>= 17 waves of CTAs at CAS(16,16))
        const size_t group = 4 * piece;
        size_t group_cur = piece;
        size_t next_launch = group_cur;
        for (size_t off = 0; off < total; off += piece) {
            const size_t cnt = std::min(piece, total - off);
            parallel_copy(h->hC + off, C_host + off, cnt);
            F2_CUDA(cudaMemcpyAsync(h->dC.p + off, h->hC + off, cnt * sizeof(double), cudaMemcpyHostToDevice,
                                    h->copy_stream));
            const size_t arrived = off + cnt;
            if (streamed && (arrived >= next_launch || arrived == total)) {
                const int64_t rows_in = arrived == total ? s.nas : static_cast<int64_t>(arrived / s.nbs);
                if (rows_in > rows_done) {
                    F2_CUDA(cudaEventRecord(h->copy_ev, h->copy_stream));
                    cudaStream_t ts = (streamed_launches & 1) ? h->own_stream2 : h->own_stream;
                    F2_CUDA(cudaStreamWaitEvent(ts, h->copy_ev, 0));
                    launch_t3g(*h, h->dC.p, h->dS.p, ts, static_cast<int>(rows_done),
                               static_cast<int>(rows_in - rows_done));
                    streamed_launches += 1;
                    rows_done = rows_in;
                }
                group_cur = std::min(group, 2 * group_cur);
                next_launch = arrived + group_cur;
            }
        }
        F2_CUDA(cudaEventRecord(h->copy_ev, h->copy_stream));
        F2_CUDA(cudaStreamWaitEvent(h->own_stream, h->copy_ev, 0));
        if (streamed) { // everything that follows on own_stream comes after the launches on the second stream
            F2_CUDA(cudaEventRecord(h->t3_ev, h->own_stream2));
            F2_CUDA(cudaStreamWaitEvent(h->own_stream, h->t3_ev, 0));
        }
        stamp("staged + H2D enqueued");
        if (trace) {
            F2_CUDA(cudaStreamSynchronize(h->copy_stream));
            stamp("H2D done");
        }
        // the way back: sigma rows travel in pieces, each followed by an event the unstaging copy waits for
        struct Piece {
            size_t off, cnt;
        };
        std::vector<Piece> out;
        auto piece_event = [&](size_t i) {
            if (h->piece_ev.size() <= i) {
                const size_t old = h->piece_ev.size();
                h->piece_ev.resize(i + 1);
                for (size_t k = old; k <= i; ++k)
                    F2_CUDA(cudaEventCreateWithFlags(&h->piece_ev[k], cudaEventDisableTiming));
            }
            return h->piece_ev[i];
        };
        auto send_back = [&](size_t off, size_t cnt, cudaStream_t st) {
            for (size_t o = off; o < off + cnt; o += piece) {
                const size_t c = std::min(piece, off + cnt - o);
                F2_CUDA(cudaMemcpyAsync(h->hS + o, h->dS.p + o, c * sizeof(double), cudaMemcpyDeviceToHost, st));
                F2_CUDA(cudaEventRecord(piece_event(out.size()), st));
                out.push_back({o, c});
            }
        };
        if (streamed && s.nb >= 1) {
            // beta same-spin term last, in about sixteen row slabs: the rows of a finished slab are final and
            // cross PCIe on the copy stream while the next slab is computed
            if (!h->host_part)
                h->host_part = make_part(*h, 0, 1);
            // (the last slab's way home -- D2H and the copy out of the pinned buffer -- is the exposed tail: 16 slabs)
            int slabs = 16;
            if (const char* e = std::getenv("F2_HOST_SLABS")) // tuning hook
                slabs = std::max(1, std::atoi(e));
            if (!h->host_part->beta_planned)
                plan_beta(*h, *h->host_part, (s.ndet + slabs - 1) / slabs);
            if (!h->pass_ev)
                F2_CUDA(cudaEventCreateWithFlags(&h->pass_ev, cudaEventDisableTiming));
            const std::function<void(const BetaPass&)> after = [&](const BetaPass& bp) {
                F2_CUDA(cudaEventRecord(h->pass_ev, h->own_stream));
                F2_CUDA(cudaStreamWaitEvent(h->copy_stream, h->pass_ev, 0));
                for (const TrBlock& t : bp.tr_host)
                    send_back(static_cast<size_t>(t.sig_off), static_cast<size_t>(t.nA) * static_cast<size_t>(t.nB),
                              h->copy_stream);
            };
            apply(*h, *h->host_part, h->dC.p, h->dS.p, nullptr, h->own_stream, true, &after);
        } else {
            apply(*h, *h->single, h->dC.p, h->dS.p, nullptr, h->own_stream, streamed);
            send_back(0, total, h->own_stream);
        }
        h->launches += streamed_launches;
        stamp("kernels + D2H enqueued");
        if (trace) {
            F2_CUDA(cudaStreamSynchronize(h->own_stream));
            stamp("kernels done");
        }
        size_t returned = 0;
        for (size_t i = 0; i < out.size(); ++i) {
            F2_CUDA(cudaEventSynchronize(h->piece_ev[i]));
            if (trace && (i == 0 || i + 1 == out.size()))
                stamp(i == 0 ? "first sigma piece on the host" : "last sigma piece on the host");
            parallel_copy(sigma_host + out[i].off, h->hS + out[i].off, out[i].cnt);
            returned += out[i].cnt;
        }
        if (returned != total)
            throw std::logic_error("f2_sigma_apply_host: sigma pieces do not cover the vector");
        F2_CUDA(cudaStreamSynchronize(h->copy_stream));
        F2_CUDA(cudaStreamSynchronize(h->own_stream));
        stamp("done");
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_times(f2_sigma* h, double* out4) {
    try {
        F2_CUDA(cudaSetDevice(h->device));
        collect_times(*h);
        for (int i = 0; i < 4; ++i)
            out4[i] = h->builds ? h->t_acc[i] / static_cast<double>(h->builds) : 0.0;
        if (h->algorithm == "Knowles-Handy") {
            // the reference's KH path has one two-electron stage: [total, H2_kh, 0, 0] (ci_sigma_builder.cc:194-199)
            out4[1] += out4[2] + out4[3];
            out4[2] = out4[3] = 0.0;
        }
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_last_ms(f2_sigma* h, double* out5) {
    try {
        F2_CUDA(cudaSetDevice(h->device));
        collect_times(*h);
        std::memcpy(out5, h->last_ms, sizeof(h->last_ms));
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int64_t f2_sigma_launch_count(const f2_sigma* h) { return h->launches; }

int f2_sigma_part_layout(f2_sigma* h, int rank, int world, int64_t* sizes2, int64_t* rows6) {
    try {
        Part& P = shard_of(*h, rank, world);
        sizes2[0] = P.local_size;
        sizes2[1] = P.y_size;
        const auto& blocks = h->S->host->blocks;
        for (size_t bi = 0; bi < blocks.size(); ++bi) {
            const auto& blk = blocks[bi];
            int64_t* o = rows6 + 6 * bi;
            o[0] = P.lo[blk.cA];
            o[1] = P.hi[blk.cA] - P.lo[blk.cA];
            o[2] = P.loc_off[bi];
            o[3] = P.cl[blk.cB];
            o[4] = P.ch[blk.cB] - P.cl[blk.cB];
            o[5] = P.y_off[bi];
        }
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_apply_part(f2_sigma* h, const double* C_full_dev, double* sigma_local_dev, double* exch_dev,
                        int rank, int world, void* stream) {
    try {
        Part& P = shard_of(*h, rank, world);
        apply(*h, P, C_full_dev, sigma_local_dev, exch_dev, static_cast<cudaStream_t>(stream));
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_rows_form(f2_sigma* h, int rank, int world, int64_t* out4) {
    try {
        if (!h || !out4)
            throw std::invalid_argument("null argument");
        out4[0] = out4[1] = out4[2] = out4[3] = 0;
        if (world > 1 && rows_form_ok(*h)) {
            Part& P = shard_of(*h, rank, world);
            if (h->t3g || P.srow_slot.p) {
                out4[0] = 1;
                out4[1] = P.slot_elems;
                out4[2] = h->S->host->nbs;
                out4[3] = P.y_size;
            }
        }
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_apply_rows(f2_sigma* h, const double* C_local_dev, const double* xcols_dev, double* sigma_local_dev,
                        double* ycols_dev, double* scatter_dev, int rank, int world, int phases, void* stream) {
    try {
        if (!h || !C_local_dev || !xcols_dev || !sigma_local_dev || !ycols_dev || !scatter_dev)
            throw std::invalid_argument("null argument");
        if (world <= 1 || !rows_form_ok(*h) || !(h->t3g || shard_of(*h, rank, world).srow_slot.p))
            throw std::runtime_error("CISigmaBuilder: the rows form of the sharded build is not available for "
                                     "this string space (f2_sigma_rows_form)");
        // t3g_kernel brings its rows of C in with TMA bulk copies; t3gp_kernel gathers and has no such need
        if (h->t3g && (reinterpret_cast<uintptr_t>(C_local_dev) & 15) != 0)
            throw std::invalid_argument("CISigmaBuilder: the local trial vector must be 16-byte aligned");
        if (phases < 1 || phases > 15)
            throw std::invalid_argument("CISigmaBuilder: phases must be a combination of 1, 2 (= 4 | 8), 4 and 8");
        Part& P = shard_of(*h, rank, world);
        apply_rows(*h, P, C_local_dev, xcols_dev, sigma_local_dev, ycols_dev, scatter_dev, phases,
                   static_cast<cudaStream_t>(stream));
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_apply_rows_t3_piece(f2_sigma* h, const double* C_local_dev, double* scatter_dev, int rank, int world,
                                 int64_t row_first, int64_t row_count, int flags, void* stream) {
    try {
        if (!h || !C_local_dev || !scatter_dev)
            throw std::invalid_argument("null argument");
        if (world <= 1 || !rows_form_ok(*h) || !h->t3g)
            throw std::runtime_error("CISigmaBuilder: the opposite-spin term in pieces needs the rows form of a space with one "
                                     "determinant block (f2_sigma_rows_form, f2_sigma_t3_shape)");
        if ((reinterpret_cast<uintptr_t>(C_local_dev) & 15) != 0)
            throw std::invalid_argument("CISigmaBuilder: the local trial vector must be 16-byte aligned");
        Part& P = shard_of(*h, rank, world);
        apply_rows_t3_piece(*h, P, C_local_dev, scatter_dev, row_first, row_count, flags, static_cast<cudaStream_t>(stream));
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_t3_shape(const f2_sigma* h, int64_t* out5) {
    out5[5] = (h->t3g || h->t3gm) ? 1 : 0;
    out5[0] = h->M;
    if (h->t3g) { // rows of A' actually multiplied, expressed per column of the GEMM K dimension of the list
        const int rows = h->t3g_nm16 * 16 + h->t3g_has8 * 8 + h->t3g_nx;
        const int kmax = std::max(h->S->host->device_lists().as_max, 1);
        out5[0] = h->M;
        out5[1] = (static_cast<int64_t>(h->t3g_mpad) * rows + kmax - 1) / kmax;
        out5[2] = kmax;
        out5[3] = h->single ? h->single->t3_kcols : 0;
        out5[4] = h->compressed ? 1 : 0;
        return F2_OK;
    }
    out5[1] = (static_cast<int64_t>(h->t3_warps) * 16 + (h->t3_tail ? 8 : 0)) * h->t3_msplit;
    out5[2] = h->S->host->device_lists().as_max;
    out5[3] = h->single ? h->single->t3_kcols : 0;
    out5[4] = h->compressed ? 1 : 0;
    return F2_OK;
}

int f2_sigma_hdiag(f2_sigma* h, double* out_dev, void* stream) {
    try {
        NvtxRange nvtx("determinant energies");
        F2_CUDA(cudaSetDevice(h->device));
        if (h->hd_rows == 0)
            return F2_OK;
        hdiag_kernel<<<h->hd_rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            out_dev, h->hd_blocks.p, h->hd_nblocks, h->ea.p, h->eb.p, h->ja.p, h->D->b_str.p,
            h->norb, h->E);
        F2_CUDA(cudaGetLastError());
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

namespace {

int64_t binom64(int n, int k) {
    if (k < 0 || k > n)
        return 0;
    int64_t x = 1;
    for (int i = 1; i <= k; ++i)
        x = x * (n - k + i) / i;
    return x;
}

f2_sigma::RdmLists& rdm_device_lists(f2_sigma& b, int spin, int level) {
    auto& d = b.rdm_lists[spin][level];
    if (!d.up) {
        const HoleLists& hl = b.S->host->hole_lists(spin, level);
        d.ptr.upload(hl.ptr);
        d.ent.upload(hl.ent);
        // cudaMemcpy from pageable memory returns once the data is staged: the DMA tail must have landed before a
        // kernel on the caller's (possibly non-blocking) stream reads the lists
        F2_CUDA(cudaDeviceSynchronize());
        d.up = true;
    }
    return d;
}

extern "C++" {
template <int MT>
void launch_rdm_gram(const RdmGramParams& q, unsigned nK, bool al16, bool same, cudaStream_t st) {
    auto go = [&](auto kernel) {
        const size_t smem = rdm_gram_smem_bytes(MT, same);
        F2_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kernel<<<nK, GRAM_THREADS, smem, st>>>(q);
    };
    if (al16)
        same ? go(rdm_gram_kernel<MT, true, true>) : go(rdm_gram_kernel<MT, true, false>);
    else
        same ? go(rdm_gram_kernel<MT, false, true>) : go(rdm_gram_kernel<MT, false, false>);
}
} // extern "C++"

// one-spin block (1-RDM, same-spin 2-RDM, same-spin 3-RDM with at most 64 parents per hole string) as Gram matrices of
// the parent rows (rdm_gram_kernel): alpha on the vectors themselves, beta on their transposed row slabs (the slabs and
// plans of the beta same-spin sigma term, inside the set_memory budget)
void rdm_gram(f2_sigma& b, int spin, int level, const double* L, const double* R, double* out, int64_t Pops,
              cudaStream_t st, int part, int nparts) {
    const CIStringsHost& hs = *b.S->host;
    const DevStrings& D = *b.D;
    const HoleLists& HL = hs.hole_lists(spin, level);
    NvtxRange nvtx("rdm: Gram form (1-RDM / same-spin blocks)");
    const unsigned nK_all = static_cast<unsigned>(HL.nK());
    const unsigned kfirst = static_cast<unsigned>(static_cast<uint64_t>(nK_all) * part / nparts);
    const unsigned nK = static_cast<unsigned>(static_cast<uint64_t>(nK_all) * (part + 1) / nparts) - kfirst;
    if (nK == 0)
        return;
    auto& dl = rdm_device_lists(b, spin, level);
    Part& P = *b.single;
    const bool same = L == R;
    RdmGramParams q;
    q.out = out;
    q.hptr = dl.ptr.p;
    q.hent = dl.ent.p;
    q.s_cls = spin == 0 ? D.a_cls.p : D.b_cls.p;
    q.s_loc = spin == 0 ? D.a_loc.p : D.b_loc.p;
    q.Pops = static_cast<int>(Pops);
    q.kfirst = kfirst;
    const int mt = std::max(1, (HL.max_parents + 15) / 16);
    auto launch = [&](const SSPlan& plan, const double* xl, const double* xr) {
        if (plan.ngroups == 0)
            return;
        q.XL = xl;
        q.XR = xr;
        q.xoff = plan.xoff.p;
        q.g_ncols = plan.g_ncols.p;
        q.g_ldx = plan.g_ldx.p;
        q.ngroups = plan.ngroups;
        const bool al16 = plan.al16 && ((reinterpret_cast<uintptr_t>(xl) | reinterpret_cast<uintptr_t>(xr)) & 15) == 0;
        switch (mt) {
        case 1: launch_rdm_gram<1>(q, nK, al16, same, st); break;
        case 2: launch_rdm_gram<2>(q, nK, al16, same, st); break;
        case 3: launch_rdm_gram<3>(q, nK, al16, same, st); break;
        case 4: launch_rdm_gram<4>(q, nK, al16, same, st); break;
        default: throw std::logic_error("rdm_gram: more than 64 parents");
        }
        b.launches += 1;
    };
    if (spin == 0) {
        launch(P.alpha_plan, L, R);
    } else {
        if (!P.beta_planned)
            plan_beta(b, P);
        for (const auto& bp : P.beta_passes) {
            if (bp->ntr_tiles == 0)
                continue;
            transpose_kernel<false><<<bp->ntr_tiles, 256, 0, st>>>(L, b.XT.p, bp->tr.p, bp->ntr, false);
            b.launches += 1;
            if (!same) {
                transpose_kernel<false><<<bp->ntr_tiles, 256, 0, st>>>(R, b.YT.p, bp->tr.p, bp->ntr, false);
                b.launches += 1;
            }
            launch(bp->plan, b.XT.p, same ? b.XT.p : b.YT.p);
        }
    }
    F2_CUDA(cudaGetLastError());
}

extern "C++" {
template <int MW, int NTP>
void launch_rdm_ab(const RdmAbParams& q, unsigned nas, size_t smem, bool long_rows, cudaStream_t st) {
    if (long_rows) {
        F2_CUDA(cudaFuncSetAttribute(rdm_ab_kernel<MW, NTP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
        rdm_ab_kernel<MW, NTP, true><<<nas, AB_THREADS, smem, st>>>(q);
    } else {
        F2_CUDA(cudaFuncSetAttribute(rdm_ab_kernel<MW, NTP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
        rdm_ab_kernel<MW, NTP, false><<<nas, AB_THREADS, smem, st>>>(q);
    }
}
} // extern "C++"

// opposite-spin 2-RDM in the gather form (rdm_ab_kernel); false when the space does not fit it (several determinant
// blocks, a row of the vector longer than shared memory, more than 19 orbitals or 128 single replacements)
bool rdm_ab(f2_sigma& b, const double* L, const double* R, double* out, cudaStream_t st, int part, int nparts) {
    NvtxRange nvtx("rdm: opposite-spin 2-RDM, gather form");
    const CIStringsHost& hs = *b.S->host;
    const DevStrings& D = *b.D;
    const DeviceLists& DL = hs.device_lists();
    if (hs.nclsA() != 1 || hs.nclsB() != 1 || hs.blocks.size() != 1 || DL.as_max > 128 || DL.as_max < 1)
        return false;
    const int n = hs.norb, n2 = n * n;
    const int mtiles = (n2 + 15) / 16, mw = (mtiles + 7) / 8;
    if (mw > 3 || hs.nbs >= (int64_t(1) << 31) - 1)
        return false;
    const int nt = (DL.as_max + 7) / 8;
    int ntp;
    if (mw == 1)
        ntp = nt <= 4 ? 4 : (nt <= 8 ? 8 : 16);
    else if (mw == 2)
        ntp = nt <= 5 ? 5 : 9;
    else
        ntp = 6;
    // rows longer than shared memory (CAS(18,18)): stage what fits beside the L stages, gather the rest from L2
    size_t smem = rdm_ab_smem_bytes(hs.nbs, ntp);
    const size_t smem_cap = size_t(227) * 1024 - 2048;
    int64_t nsm = hs.nbs;
    if (const char* e = std::getenv("F2_RDM_AB_NSM")) // test hook: exercise the long-row variant on a small space
        nsm = std::min<int64_t>(nsm, std::max<int64_t>(0, std::atoll(e)));
    if (smem > smem_cap)
        nsm = std::min<int64_t>(nsm, (static_cast<int64_t>(smem_cap) - static_cast<int64_t>(rdm_ab_smem_bytes(0, ntp))) / 8);
    if (nsm < 0)
        return false;
    const bool long_rows = nsm < hs.nbs;
    smem = rdm_ab_smem_bytes(nsm, ntp);
    const int rows = 8 * mw * 16;
    if (!b.rdm_gtab.p || b.rdm_gtab_rows != rows) {
        b.rdm_gld = (hs.nbs + 31) / 32 * 32;
        std::vector<uint32_t> fill(static_cast<size_t>(rows) * b.rdm_gld, static_cast<uint32_t>(hs.nbs));
        b.rdm_gtab.upload(fill);
        F2_CUDA(cudaDeviceSynchronize()); // the fill has landed before the table kernel on the caller's stream writes
        b.rdm_gtab_rows = rows;
        const int64_t total = hs.nbs * DL.bs_width;
        if (total > 0)
            rdm_abtab_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(D.bs_ent.p, hs.nbs, DL.bs_width, b.rdm_gld,
                                                                                        b.rdm_gtab.p);
        b.launches += 1;
    }
    RdmAbParams q;
    q.L = L;
    q.R = R;
    q.out = out;
    q.as_ptr = D.as_ptr.p;
    q.as_ent = D.as_ent.p;
    q.gtab = b.rdm_gtab.p;
    q.nbs = hs.nbs;
    q.gld = b.rdm_gld;
    q.norb = n;
    q.n2 = n2;
    q.nsm = static_cast<int>(nsm);
    q.first = static_cast<uint32_t>(hs.nas * part / nparts);
    const unsigned nas = static_cast<unsigned>(hs.nas * (part + 1) / nparts) - q.first;
    if (nas == 0)
        return true;
    switch (mw * 100 + ntp) {
    case 104: launch_rdm_ab<1, 4>(q, nas, smem, long_rows, st); break;
    case 108: launch_rdm_ab<1, 8>(q, nas, smem, long_rows, st); break;
    case 116: launch_rdm_ab<1, 16>(q, nas, smem, long_rows, st); break;
    case 205: launch_rdm_ab<2, 5>(q, nas, smem, long_rows, st); break;
    case 209: launch_rdm_ab<2, 9>(q, nas, smem, long_rows, st); break;
    case 306: launch_rdm_ab<3, 6>(q, nas, smem, long_rows, st); break;
    default: return false;
    }
    b.launches += 1;
    F2_CUDA(cudaGetLastError());
    return true;
}

// one spin block <L| (la alpha creators)(lb beta creators)(lb beta annihilators)(la alpha annihilators) |R> as the
// M x M matrix over rows opA * Pb + opB, M = C(norb, la) C(norb, lb)
// part / nparts: only the contributions of that slice of the hole strings (of the input alpha strings in the gather form):
// the sum over the parts is the block (f2_sigma_rdm_part: one part per GPU, then one all-reduce)
void rdm_block(f2_sigma& b, int la, int lb, const double* L, const double* R, double* out, cudaStream_t st,
               int part = 0, int nparts = 1) {
    if (nparts < 1 || part < 0 || part >= nparts)
        throw std::invalid_argument("f2_sigma_rdm_part: need 0 <= part < nparts");
    NvtxRange nvtx("rdm: spin block");
    const CIStringsHost& hs = *b.S->host;
    const DevStrings& D = *b.D;
    if (la < 0 || lb < 0 || la + lb < 1 || la + lb > 3)
        throw std::invalid_argument("f2_sigma_rdm_block: need 1 <= la + lb <= 3");
    const int n = hs.norb;
    const int64_t Pa = binom64(n, la), Pb = binom64(n, lb), M = Pa * Pb;
    if (M == 0)
        return;
    if (M > (int64_t(1) << 15))
        throw std::overflow_error("f2_sigma_rdm_block: more than 32768 operator tuples per side");
    F2_CUDA(cudaMemsetAsync(out, 0, static_cast<size_t>(M * M) * sizeof(double), st));
    // the reference returns zeros when a spin has too few electrons (ci_sigma_builder_1rdm.cc:15-17,
    // ci_sigma_builder_2rdm.cc:64-66,148-151, ci_sigma_builder_3rdm.cc:101-103,189-191)
    if (hs.ndet == 0 || hs.na < la || hs.nb < lb)
        return;
    const HoleLists& HA = hs.hole_lists(0, la);
    const HoleLists& HB = hs.hole_lists(1, lb);
    const int64_t nKa = static_cast<int64_t>(HA.nK()), nKb = static_cast<int64_t>(HB.nK());
    if (nKa == 0 || nKb == 0)
        return;
    // one spin only: Gram matrices of the parent rows, no D matrix (F2_RDM_GENERIC: the gather + GEMM path always)
    if ((la == 0 || lb == 0) && (la == 0 ? HB : HA).max_parents <= 64 && std::getenv("F2_RDM_GENERIC") == nullptr) {
        rdm_gram(b, la == 0 ? 1 : 0, la + lb, L, R, out, M, st, part, nparts);
        return;
    }
    if (la == 1 && lb == 1 && std::getenv("F2_RDM_GENERIC") == nullptr && rdm_ab(b, L, R, out, st, part, nparts))
        return;
    auto& LA = rdm_device_lists(b, 0, la);
    auto& LB = rdm_device_lists(b, 1, lb);
    const bool same = L == R;
    F2_CUDA(cudaFuncSetAttribute(rdm_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(rdm_gemm_smem_bytes(128))));
    F2_CUDA(cudaFuncSetAttribute(rdm_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(rdm_gemm_smem_bytes(64))));
    RdmGatherParams g;
    g.ptrA = LA.ptr.p;
    g.entA = LA.ent.p;
    g.ptrB = LB.ptr.p;
    g.entB = LB.ent.p;
    g.a_cls = D.a_cls.p;
    g.a_loc = D.a_loc.p;
    g.b_cls = D.b_cls.p;
    g.b_loc = D.b_loc.p;
    g.nB_cls = D.nB_cls.p;
    g.blk_off = D.blk_off.p;
    g.nclsB = static_cast<int>(hs.nclsB());
    g.Pb = static_cast<int>(Pb);
    g.kaList = g.kbList = nullptr;
    g.rowmap = nullptr;
    // one pass of the chunk loop: Mrows operator rows (all of them, or those of one irrep), hole strings
    // [0, nka_all) x [0, nkb_all) of the lists in g (all hole strings, or those of two matching irreps)
    auto run = [&](int64_t Mrows, int64_t nka_all, int64_t nkb_all, const int32_t* unmap) {
        if (Mrows == 0 || nka_all == 0 || nkb_all == 0)
            return;
        // the K-block budget of the reference (acquire_local_Kblock_buffers, ci_sigma_builder.cc:70-78,
        // knowles_handy.cc:15-32): two buffers of M rows inside set_memory; at least one column each
        const int64_t cap = static_cast<int64_t>(b.memory_size / ((same ? 1 : 2) * sizeof(double)));
        int64_t kc_max = cap / M; // (sized by the full row count, so that the error does not depend on the symmetry)
        kc_max -= kc_max % 2;
        if (kc_max < 2)
            throw std::runtime_error("The CI builder memory limit (" + std::to_string(b.memory_size) +
                                     " bytes) is too small for one K-block column of " + std::to_string(M) +
                                     " elements per buffer.");
        kc_max = std::min<int64_t>(cap / Mrows - (cap / Mrows) % 2, int64_t(1) << 30);
        if (const char* e = std::getenv("F2_RDM_CHUNK_COLS")) // test hook: columns per chunk (several chunks on small spaces)
            kc_max = std::max<int64_t>(2, std::min<int64_t>(kc_max, std::atoll(e) & ~int64_t(1)));
        int64_t cka, ckb; // chunk extents
        if (nkb_all <= kc_max) {
            ckb = nkb_all;
            cka = std::min<int64_t>(nka_all, kc_max / nkb_all);
        } else {
            cka = 1;
            ckb = kc_max;
        }
        ckb = std::min<int64_t>(ckb, int64_t(65535) * RDM_GATHER_THREADS);
        cka = std::min<int64_t>(cka, 65535 * int64_t(32768));
        const int64_t ld_max = (cka * ckb + 1) / 2 * 2;
        for (int v = 0; v < (same ? 1 : 2); ++v)
            if (b.rdmD[v].n < static_cast<size_t>(Mrows * ld_max))
                b.rdmD[v].alloc(static_cast<size_t>(Mrows * ld_max));
        const int bm = Mrows <= 32 ? 32 : (Mrows <= 64 ? 64 : 128);
        const int tiles = static_cast<int>((Mrows + bm - 1) / bm);
        const int64_t ka_lo = nka_all * part / nparts, ka_hi = nka_all * (part + 1) / nparts; // this part's hole strings
        for (int64_t ka0 = ka_lo; ka0 < ka_hi; ka0 += cka)
            for (int64_t kb0 = 0; kb0 < nkb_all; kb0 += ckb) {
                RdmGatherParams gg = g;
                gg.nka = static_cast<uint32_t>(std::min(cka, ka_hi - ka0));
                gg.nkb = static_cast<uint32_t>(std::min(ckb, nkb_all - kb0));
                gg.ka0 += static_cast<uint32_t>(ka0);
                gg.kb0 += static_cast<uint32_t>(kb0);
                const int64_t K = static_cast<int64_t>(gg.nka) * gg.nkb;
                gg.ldd = (K + 1) / 2 * 2;
                const dim3 grid(gg.nka, (gg.nkb + RDM_GATHER_THREADS - 1) / RDM_GATHER_THREADS);
                for (int v = 0; v < (same ? 1 : 2); ++v) {
                    gg.X = v == 0 ? L : R;
                    gg.D = b.rdmD[v].p;
                    F2_CUDA(cudaMemsetAsync(gg.D, 0, static_cast<size_t>(Mrows * gg.ldd) * sizeof(double), st));
                    rdm_gather_kernel<<<grid, RDM_GATHER_THREADS, 0, st>>>(gg);
                    b.launches += 1;
                }
                RdmGemmParams q;
                q.A = b.rdmD[0].p;
                q.B = same ? b.rdmD[0].p : b.rdmD[1].p;
                q.C = out;
                q.M = static_cast<int>(Mrows);
                q.K = K;
                q.ld = gg.ldd;
                q.unmap = unmap;
                q.ldc = static_cast<int>(M);
                const int64_t nsteps = (K + RG_BK - 1) / RG_BK;
                q.ksplit = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(nsteps, (2 * 148 + tiles * tiles - 1) / (tiles * tiles))));
                const unsigned ctas = static_cast<unsigned>(tiles * tiles * q.ksplit);
                if (bm == 32)
                    rdm_gemm_kernel<32><<<ctas, RG_THREADS, rdm_gemm_smem_bytes(32), st>>>(q);
                else if (bm == 64)
                    rdm_gemm_kernel<64><<<ctas, RG_THREADS, rdm_gemm_smem_bytes(64), st>>>(q);
                else
                    rdm_gemm_kernel<128><<<ctas, RG_THREADS, rdm_gemm_smem_bytes(128), st>>>(q);
                b.launches += 1;
            }
    };
    int nirrep = 1; // irrep labels combine by XOR: group by the power of two that covers them
    while (nirrep < hs.nirrep)
        nirrep <<= 1;
    if (nirrep <= 1 || std::getenv("F2_RDM_NO_IRREP_BLOCKS") != nullptr) {
        g.ka0 = g.kb0 = 0;
        run(M, nKa, nKb, nullptr);
        F2_CUDA(cudaGetLastError());
        return;
    }
    // ---- point-group symmetry: one pass per (irrep h of the operator tuples, irrep of Ka); see RdmGatherParams
    auto tuple_irreps = [&](int level) { // irrep of every decreasing orbital tuple, in tuple-index order
        std::vector<int> out;
        const int n = hs.norb;
        if (level == 0)
            out.push_back(0);
        else if (level == 1)
            for (int p = 0; p < n; ++p)
                out.push_back(hs.symmetry_of(1ull << p));
        else if (level == 2)
            for (int p = 0; p < n; ++p)
                for (int q = 0; q < p; ++q)
                    out.push_back(hs.symmetry_of((1ull << p) | (1ull << q)));
        else
            for (int p = 0; p < n; ++p)
                for (int q = 0; q < p; ++q)
                    for (int r = 0; r < q; ++r)
                        out.push_back(hs.symmetry_of((1ull << p) | (1ull << q) | (1ull << r)));
        return out;
    };
    const std::vector<int> irA = tuple_irreps(la), irB = tuple_irreps(lb);
    if (static_cast<int64_t>(irA.size()) != Pa || static_cast<int64_t>(irB.size()) != Pb)
        throw std::logic_error("rdm_block: operator tuple count");
    std::vector<int32_t> rowmap(static_cast<size_t>(nirrep) * M, -1), unmap(static_cast<size_t>(nirrep) * M, 0);
    std::vector<int64_t> Mh(nirrep, 0);
    for (int64_t a = 0; a < Pa; ++a)
        for (int64_t bb = 0; bb < Pb; ++bb) {
            const int h = irA[a] ^ irB[bb];
            const int64_t row = a * Pb + bb;
            rowmap[h * M + row] = static_cast<int32_t>(Mh[h]);
            unmap[h * M + Mh[h]] = static_cast<int32_t>(row);
            Mh[h] += 1;
        }
    auto by_irrep = [&](const HoleLists& HL, std::vector<uint32_t>& flat, std::vector<int64_t>& first) {
        std::vector<std::vector<uint32_t>> lists(nirrep);
        for (size_t K = 0; K < HL.nK(); ++K)
            lists[hs.symmetry_of(HL.kword[K])].push_back(static_cast<uint32_t>(K));
        first.assign(nirrep + 1, 0);
        for (int h = 0; h < nirrep; ++h) {
            first[h + 1] = first[h] + static_cast<int64_t>(lists[h].size());
            flat.insert(flat.end(), lists[h].begin(), lists[h].end());
        }
    };
    std::vector<uint32_t> kaFlat, kbFlat;
    std::vector<int64_t> kaFirst, kbFirst;
    by_irrep(HA, kaFlat, kaFirst);
    by_irrep(HB, kbFlat, kbFirst);
    DevBuf<uint32_t> dKa, dKb;
    DevBuf<int32_t> dRowmap, dUnmap;
    dKa.upload(kaFlat);
    dKb.upload(kbFlat);
    dRowmap.upload(rowmap);
    dUnmap.upload(unmap);
    F2_CUDA(cudaDeviceSynchronize()); // the uploads (pageable memory) have landed before the caller's stream reads them
    g.kaList = dKa.p;
    g.kbList = dKb.p;
    for (int h = 0; h < nirrep; ++h) {
        if (Mh[h] == 0)
            continue;
        g.rowmap = dRowmap.p + static_cast<size_t>(h) * M;
        for (int a = 0; a < nirrep; ++a) {
            const int bsym = hs.symmetry ^ h ^ a; // irrep(Ka) x irrep(Kb) = h x symmetry of the state
            if (bsym < 0 || bsym >= nirrep)
                continue;
            g.ka0 = static_cast<uint32_t>(kaFirst[a]);
            g.kb0 = static_cast<uint32_t>(kbFirst[bsym]);
            run(Mh[h], kaFirst[a + 1] - kaFirst[a], kbFirst[bsym + 1] - kbFirst[bsym], dUnmap.p + static_cast<size_t>(h) * M);
        }
    }
    F2_CUDA(cudaGetLastError());
    F2_CUDA(cudaStreamSynchronize(st)); // the index tables are freed on return
}

// segments of consecutive strings of right class c that are consecutive strings of one left class
struct ClassSegment {
    uint32_t r0, cl, l0, len;
};

std::vector<std::vector<ClassSegment>> match_classes(const StringSpace& Rs, const StringSpace& Ls) {
    std::vector<std::vector<ClassSegment>> out(Rs.nclasses());
    for (size_t c = 0; c < Rs.nclasses(); ++c) {
        auto& segs = out[c];
        for (uint32_t j = 0; j < Rs.cls[c].size(); ++j) {
            uint32_t add, cl;
            if (!Ls.find(Rs.cls[c][j], add, cl))
                continue;
            if (!segs.empty() && segs.back().cl == cl && segs.back().r0 + segs.back().len == j &&
                segs.back().l0 + segs.back().len == add)
                segs.back().len += 1;
            else
                segs.push_back({j, cl, add, 1});
        }
    }
    return out;
}

} // namespace

int f2_sigma_rdm_block(f2_sigma* h, int la, int lb, const double* L_dev, const double* R_dev, double* out_dev,
                       void* stream) {
    try {
        if (!h || !L_dev || !R_dev || !out_dev)
            throw std::invalid_argument("null argument");
        F2_CUDA(cudaSetDevice(h->device));
        rdm_block(*h, la, lb, L_dev, R_dev, out_dev, static_cast<cudaStream_t>(stream));
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_rdm_part(f2_sigma* h, int kind, const double* L_dev, const double* R_dev, double* out_dev, int part, int nparts,
                      void* stream) {
    static const int levels[][2] = {{1, 0}, {0, 1}, {-1, -1}, {2, 0}, {0, 2}, {1, 1}, {-1, -1}, {-1, -1},
                                    {3, 0}, {2, 1}, {1, 2}, {0, 3}};
    try {
        if (!h || !L_dev || !R_dev || !out_dev)
            throw std::invalid_argument("null argument");
        if (kind < 0 || kind > 11 || levels[kind][0] < 0)
            throw std::invalid_argument("f2_sigma_rdm_part: kind must be 0, 1, 3, 4, 5 or 8 .. 11 (f2_sigma_rdm)");
        F2_CUDA(cudaSetDevice(h->device));
        rdm_block(*h, levels[kind][0], levels[kind][1], L_dev, R_dev, out_dev, static_cast<cudaStream_t>(stream), part, nparts);
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_sigma_rdm(f2_sigma* h, int kind, const double* L_dev, const double* R_dev, double* out_dev, void* stream) {
    static const int levels[][2] = {{1, 0}, {0, 1}, {-1, -1}, {2, 0}, {0, 2}, {1, 1}, {-1, -1}, {-1, -1},
                                    {3, 0}, {2, 1}, {1, 2}, {0, 3}};
    if (kind < 0 || kind > 11 || levels[kind][0] < 0) {
        g_last_error = "f2_sigma_rdm: kind must be 0 (a_1rdm), 1 (b_1rdm), 3 (aa_2rdm), 4 (bb_2rdm), 5 (ab_2rdm), "
                       "8 (aaa_3rdm), 9 (aab_3rdm), 10 (abb_3rdm) or 11 (bbb_3rdm)";
        return F2_EINVAL;
    }
    return f2_sigma_rdm_block(h, levels[kind][0], levels[kind][1], L_dev, R_dev, out_dev, stream);
}

int f2_sigma_trdm1(f2_sigma* hl, f2_sigma* hr, int spin, const double* L_dev, const double* R_dev, double* out_dev,
                   void* stream) {
    try {
        NvtxRange nvtx("transition 1-RDM");
        if (!hl || !hr || !L_dev || !R_dev || !out_dev || spin < 0 || spin > 1)
            throw std::invalid_argument("f2_sigma_trdm1: null argument or spin outside {0, 1}");
        const CIStringsHost& SL = *hl->S->host;
        const CIStringsHost& SR = *hr->S->host;
        // the reference's checks and messages (ci_sigma_builder_1trdm.cc:24-34)
        if (SL.norb != SR.norb)
            throw std::runtime_error("FCI transition RDMs: The number of MOs must be the same in the two wave functions.");
        if (SL.na != SR.na || SL.nb != SR.nb)
            throw std::runtime_error("FCI transition RDMs: The number of alpha and beta electrons must be the same in the "
                                     "two wave functions.");
        if (hl->device != hr->device)
            throw std::invalid_argument("f2_sigma_trdm1: both builders must live on the same device");
        F2_CUDA(cudaSetDevice(hl->device));
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int n = SL.norb;
        F2_CUDA(cudaMemsetAsync(out_dev, 0, static_cast<size_t>(n) * n * sizeof(double), st));
        if ((spin == 0 ? SL.na : SL.nb) < 1 || SL.ndet == 0 || SR.ndet == 0)
            return F2_OK;
        // acted-on spin: strings of the right space and their single replacements found in the left space;
        // untouched spin: class-to-class segments of identical strings
        const StringSpace& Ract = SR.space(spin, K_STR_N);
        const StringSpace& Lact = SL.space(spin, K_STR_N);
        const auto segs = match_classes(SR.space(1 - spin, K_STR_N), SL.space(1 - spin, K_STR_N));
        const size_t nclsB_L = SL.nclsB(), nclsB_R = SR.nclsB();
        std::vector<TrdmEntry> ent;
        for (const auto& blkR : SR.blocks) {
            if (blkR.nA * blkR.nB == 0)
                continue;
            const int cActR = spin == 0 ? blkR.cA : blkR.cB;  // class of the acted-on strings in this block
            const int cOthR = spin == 0 ? blkR.cB : blkR.cA;
            for (uint32_t x = 0; x < Ract.cls[cActR].size(); ++x) {
                const uint64_t w = Ract.cls[cActR][x];
                for (int q = 0; q < n; ++q) {
                    if (!((w >> q) & 1))
                        continue;
                    const int sq = slater_sign(w, q);
                    const uint64_t w1 = w & ~(1ull << q);
                    for (int pp = 0; pp < n; ++pp) {
                        if ((w1 >> pp) & 1)
                            continue;
                        uint32_t addL, clsL;
                        if (!Lact.find(w1 | (1ull << pp), addL, clsL))
                            continue;
                        const int sign = sq * slater_sign(w1, pp);
                        for (const ClassSegment& sg : segs[cOthR]) {
                            const int64_t bi = spin == 0 ? SL.block_of[clsL * nclsB_L + sg.cl]
                                                         : SL.block_of[sg.cl * nclsB_L + clsL];
                            if (bi < 0 || SL.blocks[bi].nA * SL.blocks[bi].nB == 0)
                                continue;
                            const auto& blkL = SL.blocks[bi];
                            TrdmEntry e;
                            if (spin == 0) { // rows are the acted-on strings, the dot runs along a row
                                e.offL = blkL.offset + static_cast<int64_t>(addL) * blkL.nB + sg.l0;
                                e.offR = blkR.offset + static_cast<int64_t>(x) * blkR.nB + sg.r0;
                                e.strideL = e.strideR = 1;
                            } else { // columns are the acted-on strings, the dot runs down a column
                                e.offL = blkL.offset + static_cast<int64_t>(sg.l0) * blkL.nB + addL;
                                e.offR = blkR.offset + static_cast<int64_t>(sg.r0) * blkR.nB + x;
                                e.strideL = static_cast<int32_t>(blkL.nB);
                                e.strideR = static_cast<int32_t>(blkR.nB);
                            }
                            e.len = static_cast<int32_t>(sg.len);
                            e.out = (pp * n + q) | (sign < 0 ? static_cast<int32_t>(0x80000000u) : 0);
                            ent.push_back(e);
                        }
                    }
                }
            }
        }
        (void)nclsB_R;
        if (ent.empty())
            return F2_OK;
        DevBuf<TrdmEntry> dev;
        dev.upload(ent);
        const int64_t nent = static_cast<int64_t>(ent.size());
        const int64_t ctas = (nent * 32 + 255) / 256;
        trdm_kernel<<<static_cast<unsigned>(ctas), 256, 0, st>>>(L_dev, R_dev, dev.p, nent, out_dev);
        hl->launches += 1;
        F2_CUDA(cudaGetLastError());
        F2_CUDA(cudaStreamSynchronize(st)); // the entry table is freed on return
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

static int hdiag_part_impl(f2_sigma* h, double* out_local_dev, int rank, int world, bool spin_averaged, void* stream);

int f2_sigma_hdiag_part(f2_sigma* h, double* out_local_dev, int rank, int world, void* stream) {
    return hdiag_part_impl(h, out_local_dev, rank, world, false, stream);
}

int f2_sigma_hdiag_spin_averaged(f2_sigma* h, double* out_local_dev, int rank, int world, void* stream) {
    return hdiag_part_impl(h, out_local_dev, rank, world, true, stream);
}

static int hdiag_part_impl(f2_sigma* h, double* out_local_dev, int rank, int world, bool spin_averaged, void* stream) {
    try {
        if (!h || !out_local_dev)
            throw std::invalid_argument("null argument");
        F2_CUDA(cudaSetDevice(h->device));
        Part& P = shard_of(*h, rank, world);
        const CIStringsHost& hs = *h->S->host;
        std::vector<HdiagBlock> hb;
        uint32_t tile0 = 0;
        for (size_t bi = 0; bi < hs.blocks.size(); ++bi) {
            const auto& blk = hs.blocks[bi];
            const int64_t rows = P.hi[blk.cA] - P.lo[blk.cA];
            if (blk.nA * blk.nB == 0 || rows == 0)
                continue;
            HdiagBlock x;
            x.off = world == 1 ? blk.offset : P.loc_off[bi];
            x.a0 = hs.space(0, K_STR_N).cls_off[blk.cA] + static_cast<uint32_t>(P.lo[blk.cA]);
            x.b0 = hs.space(1, K_STR_N).cls_off[blk.cB];
            x.nA = static_cast<int32_t>(rows);
            x.nB = static_cast<int32_t>(blk.nB);
            x.tile0 = tile0;
            tile0 += static_cast<uint32_t>(rows);
            hb.push_back(x);
        }
        if (tile0 == 0)
            return F2_OK;
        DevBuf<HdiagBlock> dev;
        dev.upload(hb);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        hdiag_kernel<<<tile0, 256, 0, st>>>(out_local_dev, dev.p, static_cast<int>(hb.size()), h->ea.p, h->eb.p,
                                            h->ja.p, h->D->b_str.p, h->norb, h->E, h->D->a_str.p,
                                            spin_averaged ? SpinAvg{h->kx.p, h->psame.p} : SpinAvg{nullptr, nullptr});
        F2_CUDA(cudaGetLastError());
        F2_CUDA(cudaStreamSynchronize(st)); // the block table is freed on return
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

} // extern "C"

#include "dist_nccl.inc"
