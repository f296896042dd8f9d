// C ABI of the engine (include/clusttraj_b200.h): device buffers, mode dispatch, and the chunked
// compute / device->host streaming pipeline.  No CPU fallback anywhere: every entry point either runs
// the CUDA kernels or returns an error.
#include "../../include/clusttraj_b200.h"
#include "common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace ct {
cudaError_t launch_pack(const double* raw, const int* keep, int M, int N, int K, int ncen, int m_pad, int kq_total,
                        double* centred, double* G, double* packed, int sm_count, cudaStream_t stream);
cudaError_t launch_gram(const GramParams& p, int variant, int sm_count, cudaStream_t stream);
int gram_row_align(int variant);
// int8 tensor-core (tcgen05) variant of the Gram kernel, ozaki.cu
size_t oz_packA_bytes(int M, int K);
size_t oz_packB_bytes(int M, int K);
size_t oz_g_entries(int M);
int oz_row_align();
int oz_kernel_mode(int K);  // 4 / 2: A operand in TMEM (column tiles of 32 / 16 frames); 0: A through shared memory
cudaError_t launch_quant_pack(const double* centred, int M, int K, uint8_t* packA, uint8_t* packB, uint8_t* packT, double* Gq,
                              double* qinfo, unsigned long long* maxbits, int sm_count, cudaStream_t stream);
cudaError_t launch_gram_i8(OzParams p, int K, int sm_count, cudaStream_t stream);
cudaError_t launch_fixup(const double* centred, int M, int K, const FixList& fix, double* out, long long idx_base,
                         long long n_entries, unsigned long long* flagged_total, int sm_count, cudaStream_t stream);
cudaError_t launch_consumers(const double* d, long long L, int M, const int* labels, int n_clusters, double* scratch,
                             unsigned long long* best_bits, long long* medoid, int sm_count, cudaStream_t stream);
int consumers_scratch_doubles(int M);
int consumers_sum_offset(int M);
size_t linkage_scratch_bytes(int M, int sm_count);
cudaError_t run_linkage(const double* cond, double* work, size_t work_doubles, int M, int method, unsigned char* scratch,
                        int sm_count, double* zraw_host, int* done_flag, long long* steps_out, cudaStream_t stream);
void linkage_finish(const double* zraw, int M, double* Z);
struct PairPlan;
struct PairPlanHost;
int pair_plan_build(PairPlanHost** out, const int32_t* kept_z, int K, int natoms, const ct_opts* o, std::string& err);
void pair_plan_free(PairPlanHost* p);
int pair_plan_frame_elements(PairPlanHost* h, const int32_t* z, int M, int K, std::string& err);
cudaError_t launch_pairs_rows(const PairPlanHost* plan, const double* centred, int M, int K, int row_begin, int row_end,
                              double* out, long long idx_base, unsigned long long* counters, int sm_count,
                              cudaStream_t stream);
cudaError_t launch_pairs_list(const PairPlanHost* plan, const double* centred, int M, int K, const long long* pairs_dev,
                              long long n_pairs, double* values_dev, int* perms_dev, unsigned long long* counters,
                              int sm_count, cudaStream_t stream, double* xf_dev);
}  // namespace ct

struct ct_engine {
    int device = 0;
    int sm_count = 0;
    std::string err;
    cudaStream_t s_compute = nullptr, s_copy = nullptr, s_peer = nullptr;
    // frames
    int64_t M = 0;
    int N = 0, K = 0, natoms = 0, m_pad = 0, kq_total = 0;
    bool loaded = false;
    bool per_frame_elements = false;   // ct_set_frame_elements replaced the plan's index lists (reset by ct_load_frames)
    bool use_gram = false;  // pure Kabsch modes go through the DMMA Gram kernel
    ct_opts opts{};
    std::vector<int32_t> excl;
    double* d_raw = nullptr; size_t raw_cap = 0;
    int* d_keep = nullptr; size_t keep_cap = 0;
    double* d_centred = nullptr; size_t centred_cap = 0;
    double* d_G = nullptr; size_t g_cap = 0;
    double* d_packed = nullptr; size_t packed_cap = 0;
    // int8 tensor-core path (ozaki.cu): digit slices in both operand layouts, G of the quantised coordinates
    bool use_i8 = false;
    uint8_t* d_packA = nullptr; size_t packA_cap = 0;
    uint8_t* d_packB = nullptr; size_t packB_cap = 0;
    uint8_t* d_packT = nullptr; size_t packT_cap = 0;
    double* d_Gq = nullptr; size_t gq_cap = 0;
    double* d_qinfo = nullptr;               // [0] 2^-2f, [1] f, [2] max |x|, [4..11] the magic constant of the int -> double trick
    unsigned long long* d_maxbits = nullptr;
    double quant_f = 0.0;
    ct::PairPlanHost* plan = nullptr;
    // outputs
    double* d_slab = nullptr; size_t slab_cap = 0;       // whole-slab device buffer (rows_device)
    int64_t slab_full_M = 0;                             // > 0: d_slab holds the WHOLE matrix of that many frames
    // consumers of the matrix (ct_matrix_consumers)
    int* d_labels = nullptr; size_t labels_cap = 0;
    double* d_cons = nullptr; size_t cons_cap = 0;
    unsigned long long* d_best = nullptr; size_t best_cap = 0;
    long long* d_medoid = nullptr; size_t medoid_cap = 0;
    // staging ring for copies into pageable host memory (ct_copy_slab)
    double* h_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    // hierarchical clustering (ct_linkage)
    double* d_link = nullptr; size_t link_cap = 0;           // working copy of the matrix the NN-chain methods update
    unsigned char* d_link_scratch = nullptr; size_t link_scratch_cap = 0;
    int* h_link_flag = nullptr;                              // pinned
    double* d_chunk[2] = {nullptr, nullptr}; size_t chunk_cap = 0;  // streaming double buffer
    int2* d_fix_pairs = nullptr; unsigned int* d_fix_count = nullptr; unsigned int fix_cap = 1u << 20;
    unsigned long long* d_counters = nullptr;  // [0] flagged pairs, [1] lsap steps
    cudaEvent_t ev_chunk_done[2] = {nullptr, nullptr}, ev_copy_done[2] = {nullptr, nullptr}, ev_peer_done[2] = {nullptr, nullptr};
    int64_t reserved_M = 0;                                  // > 0: d_slab has room for the whole matrix of that many frames
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr;
    ct_stats stats{};
    int gram_variant = -1;  // -1 = auto: int8 tensor-core kernel when the coordinates quantise finely enough, else
                            // the DMMA kernel (variant by atom count); -2 = DMMA, variant by atom count; 0..3 = force a DMMA
                            // variant; 10 = force int8
    int min_quant_f = 22;   // auto mode needs a quantum <= 2^-22 A (|x| < 256 A): worst-case |dRMSD| 4e-7 A
    size_t chunk_entries = (size_t)32 << 20;  // 256 MiB of float64 per streaming chunk
    double flag_rel = 1e-9;
};

static thread_local std::string g_err;

namespace ct {
void set_global_error(const std::string& msg) { g_err = msg; }   // for the host-only sources (xyzio.cu)
}  // namespace ct

#define CT_CUDA(e, call)                                                                                   \
    do {                                                                                                    \
        cudaError_t _c = (call);                                                                            \
        if (_c != cudaSuccess) {                                                                            \
            char _b[512];                                                                                   \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_c), __FILE__, __LINE__); \
            if (e) (e)->err = _b; else g_err = _b;                                                          \
            return CT_ERR_CUDA;                                                                             \
        }                                                                                                   \
    } while (0)

static int fail(ct_engine* e, int code, const std::string& msg) {
    if (e) e->err = msg; else g_err = msg;
    return code;
}

template <class T>
static int ensure(ct_engine* e, T*& ptr, size_t& cap, size_t need) {
    if (need <= cap && ptr) return CT_OK;
    if (ptr) { CT_CUDA(e, cudaFree(ptr)); ptr = nullptr; cap = 0; }
    CT_CUDA(e, cudaMalloc(&ptr, need * sizeof(T)));
    cap = need;
    return CT_OK;
}

// scipy.cluster.hierarchy.linkage refuses NaN / infinite distances with a ValueError; the medoid reduction orders sums by
// their bit patterns, which needs distances >= 0.  One pass over the resident matrix: flag[0] = entries that are not
// finite, flag[1] = negative entries.
__global__ void __launch_bounds__(256) matrix_check_kernel(const double* __restrict__ d, long long n, unsigned long long* flag) {
    unsigned long long bad = 0, neg = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const double v = d[k];
        bad += !(fabs(v) <= 1.7976931348623157e308);
        neg += v < 0.0;
    }
    bad = __reduce_add_sync(0xffffffffu, (unsigned)bad);
    neg = __reduce_add_sync(0xffffffffu, (unsigned)neg);
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(flag, bad);
        if (neg) atomicAdd(flag + 1, neg);
    }
}

// CT_ERR_ARG with SciPy's wording when the matrix resident in e->d_slab holds non-finite (or, for the medoids, negative) values
static int check_matrix_values(ct_engine* e, long long L, bool refuse_negative, const char* who) {
    CT_CUDA(e, cudaMemsetAsync(e->d_counters + 2, 0, sizeof(unsigned long long) * 2, e->s_compute));
    matrix_check_kernel<<<e->sm_count * 8, 256, 0, e->s_compute>>>(e->d_slab, L, e->d_counters + 2);
    CT_CUDA(e, cudaGetLastError());
    unsigned long long c[2];
    CT_CUDA(e, cudaMemcpyAsync(c, e->d_counters + 2, sizeof c, cudaMemcpyDeviceToHost, e->s_compute));
    CT_CUDA(e, cudaStreamSynchronize(e->s_compute));
    if (c[0]) return fail(e, CT_ERR_ARG, std::string(who) + ": The condensed distance matrix must contain only finite values.");
    if (c[1] && refuse_negative)
        return fail(e, CT_ERR_ARG, std::string(who) + ": the condensed distance matrix must not contain negative distances.");
    return CT_OK;
}

extern "C" {

const char* ct_version(void) { return "clusttraj_b200 0.1 (sm_100a)"; }

int ct_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* ct_last_error(const ct_engine* e) { return e ? e->err.c_str() : g_err.c_str(); }

static int engine_init(ct_engine* e, int device) {
    e->device = device;
    CT_CUDA(e, cudaSetDevice(device));
    cudaDeviceProp prop;
    CT_CUDA(e, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(e, CT_ERR_UNSUPPORTED, "clusttraj_b200 kernels are built for sm_100a (B200) only");
    e->sm_count = prop.multiProcessorCount;
    CT_CUDA(e, cudaStreamCreateWithFlags(&e->s_compute, cudaStreamNonBlocking));
    CT_CUDA(e, cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking));
    CT_CUDA(e, cudaStreamCreateWithFlags(&e->s_peer, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        CT_CUDA(e, cudaEventCreateWithFlags(&e->ev_chunk_done[k], cudaEventDisableTiming));
        CT_CUDA(e, cudaEventCreateWithFlags(&e->ev_copy_done[k], cudaEventDisableTiming));
        CT_CUDA(e, cudaEventCreateWithFlags(&e->ev_peer_done[k], cudaEventDisableTiming));
    }
    CT_CUDA(e, cudaEventCreate(&e->ev_a));
    CT_CUDA(e, cudaEventCreate(&e->ev_b));
    CT_CUDA(e, cudaEventCreate(&e->ev_c));
    CT_CUDA(e, cudaEventCreate(&e->ev_d));
    CT_CUDA(e, cudaMalloc(&e->d_fix_pairs, sizeof(int2) * e->fix_cap));
    CT_CUDA(e, cudaMalloc(&e->d_fix_count, sizeof(unsigned int)));
    CT_CUDA(e, cudaMalloc(&e->d_counters, sizeof(unsigned long long) * 4));
    CT_CUDA(e, cudaMalloc(&e->d_qinfo, sizeof(double) * 16));
    CT_CUDA(e, cudaMalloc(&e->d_maxbits, sizeof(unsigned long long)));
    return CT_OK;
}

int ct_engine_create(int device, ct_engine** out) {
    if (!out) return fail(nullptr, CT_ERR_ARG, "ct_engine_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t c = cudaGetDeviceCount(&n);
    if (c != cudaSuccess || n == 0)
        return fail(nullptr, CT_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(c) +
                                              " (clusttraj_b200 has no CPU fallback)");
    if (device < 0 || device >= n) return fail(nullptr, CT_ERR_ARG, "ct_engine_create: bad device ordinal");
    ct_engine* e = new ct_engine();
    const int rc = engine_init(e, device);
    if (rc != CT_OK) {
        g_err = e->err;            // the handle does not survive: the message goes to ct_last_error(NULL)
        ct_engine_destroy(e);      // frees whatever was created before the failure (every member starts out NULL)
        return rc;
    }
    const char* v = getenv("CLUSTTRAJ_B200_GRAM_VARIANT");
    if (v) e->gram_variant = atoi(v);
    const char* mq = getenv("CLUSTTRAJ_B200_MIN_QUANT_BITS");  // test hook: force the fall-back decision
    if (mq) e->min_quant_f = atoi(mq);
    const char* ce = getenv("CLUSTTRAJ_B200_CHUNK_ENTRIES");
    if (ce && atoll(ce) > 0) e->chunk_entries = (size_t)atoll(ce);
    const char* fc = getenv("CLUSTTRAJ_B200_FIX_CAPACITY");  // test hook: force the list-overflow path
    if (fc && atoll(fc) > 0 && (unsigned long long)atoll(fc) < e->fix_cap) e->fix_cap = (unsigned int)atoll(fc);
    *out = e;
    return CT_OK;
}

void ct_engine_destroy(ct_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    cudaFree(e->d_raw); cudaFree(e->d_keep); cudaFree(e->d_centred); cudaFree(e->d_G); cudaFree(e->d_packed);
    cudaFree(e->d_packA); cudaFree(e->d_packB); cudaFree(e->d_packT); cudaFree(e->d_Gq); cudaFree(e->d_qinfo); cudaFree(e->d_maxbits);
    cudaFree(e->d_labels); cudaFree(e->d_cons); cudaFree(e->d_best); cudaFree(e->d_medoid);
    cudaFree(e->d_link); cudaFree(e->d_link_scratch);
    for (int k = 0; k < 2; ++k) {
        if (e->h_stage[k]) cudaFreeHost(e->h_stage[k]);
        if (e->ev_stage[k]) cudaEventDestroy(e->ev_stage[k]);
    }
    if (e->h_link_flag) cudaFreeHost(e->h_link_flag);
    cudaFree(e->d_slab); cudaFree(e->d_chunk[0]); cudaFree(e->d_chunk[1]);
    cudaFree(e->d_fix_pairs); cudaFree(e->d_fix_count); cudaFree(e->d_counters);
    for (int k = 0; k < 2; ++k) {
        if (e->ev_chunk_done[k]) cudaEventDestroy(e->ev_chunk_done[k]);
        if (e->ev_copy_done[k]) cudaEventDestroy(e->ev_copy_done[k]);
        if (e->ev_peer_done[k]) cudaEventDestroy(e->ev_peer_done[k]);
    }
    for (cudaEvent_t ev : {e->ev_a, e->ev_b, e->ev_c, e->ev_d}) if (ev) cudaEventDestroy(ev);
    for (cudaStream_t st : {e->s_compute, e->s_copy, e->s_peer}) if (st) cudaStreamDestroy(st);
    if (e->plan) ct::pair_plan_free(e->plan);
    cudaGetLastError();   // a partially built engine leaves nothing sticky behind
    delete e;
}

int ct_host_alloc(void** ptr, uint64_t bytes) {
    if (!ptr) return fail(nullptr, CT_ERR_ARG, "ct_host_alloc: ptr is NULL");
    CT_CUDA((ct_engine*)nullptr, cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return CT_OK;
}

int ct_host_free(void* ptr) {
    if (ptr) CT_CUDA((ct_engine*)nullptr, cudaFreeHost(ptr));
    return CT_OK;
}

int64_t ct_rows_pairs(int64_t M, int64_t row_begin, int64_t row_end) {
    if (M < 2) return 0;
    if (row_begin < 0) row_begin = 0;
    if (row_end > M) row_end = M;
    if (row_end <= row_begin) return 0;
    const int64_t last = std::min<int64_t>(row_end, M - 1);
    if (last <= row_begin) return 0;
    return ct::condensed_row_base(M, last) - ct::condensed_row_base(M, row_begin);
}

int ct_load_frames(ct_engine* e, const double* coords, const int32_t* atomic_nums, int64_t M, int32_t N,
                   const ct_opts* opts) {
    if (!e) return fail(nullptr, CT_ERR_ARG, "ct_load_frames: engine is NULL");
    if (!coords || !atomic_nums || !opts) return fail(e, CT_ERR_ARG, "ct_load_frames: NULL argument");
    if (M < 1 || N < 1) return fail(e, CT_ERR_ARG, "ct_load_frames: need M >= 1 frames and N >= 1 atoms");
    if (M > (int64_t)1 << 30) return fail(e, CT_ERR_ARG, "ct_load_frames: too many frames");
    if (opts->n_excl < 0 || (opts->n_excl > 0 && !opts->reorder_excl))
        return fail(e, CT_ERR_ARG, "ct_load_frames: bad reorder_excl");
    if (opts->reorder_mode < 0 || opts->reorder_mode > 2)
        return fail(e, CT_ERR_UNSUPPORTED, "only reorder_mode 0 (none), 1 (hungarian) and 2 (distance) are implemented");
    if (opts->solute_natoms < 0 || opts->solute_natoms > N)
        return fail(e, CT_ERR_ARG, "ct_load_frames: solute_natoms out of range");
    CT_CUDA(e, cudaSetDevice(e->device));
    e->loaded = false;
    e->per_frame_elements = false;
    e->opts = *opts;
    e->excl.assign(opts->reorder_excl, opts->reorder_excl + opts->n_excl);
    e->opts.reorder_excl = e->excl.data();

    // kept atoms and the centring window (distmat.py:116-118, 134-165)
    std::vector<int> keep;
    std::vector<int32_t> kept_z;
    for (int k = 0; k < N; ++k)
        if (!opts->no_hydrogen || atomic_nums[k] != 1) { keep.push_back(k); kept_z.push_back(atomic_nums[k]); }
    const int K = (int)keep.size();
    if (K == 0) return fail(e, CT_ERR_ARG, "ct_load_frames: the hydrogen mask leaves no atoms");
    int natoms = opts->solute_natoms;
    if (opts->no_hydrogen) {
        natoms = 0;
        for (int k = 0; k < opts->solute_natoms; ++k) natoms += atomic_nums[k] != 1;
    }
    if (opts->solute_natoms && natoms == 0) return fail(e, CT_ERR_ARG, "ct_load_frames: the solute has no kept atoms");
    for (int32_t x : e->excl)
        if (x < 0 || x >= K) return fail(e, CT_ERR_ARG, "ct_load_frames: reorder_excl entry outside the kept atoms");
    const int ncen = opts->solute_natoms ? natoms : K;
    // Pure Kabsch value (dispatch rows 1-2 of SURVEY §8a): no reorder and no solute weighting.
    e->use_gram = (opts->reorder_mode == 0) && !(opts->solute_natoms && opts->weight_solute != 0.0);
    if (!opts->solute_natoms && opts->weight_solute != 0.0)
        return fail(e, CT_ERR_UNSUPPORTED, "weight_solute without solute_natoms divides by zero in the reference");

    e->M = M; e->N = N; e->K = K; e->natoms = natoms;
    e->m_pad = (int)((M + 63) / 64 * 64);
    const int k_pad = (K + 3) / 4 * 4;  // the Gram kernel handles a half-filled last stage
    e->kq_total = k_pad / 4;

    int rc;
    if ((rc = ensure(e, e->d_raw, e->raw_cap, (size_t)M * N * 3))) return rc;
    if ((rc = ensure(e, e->d_keep, e->keep_cap, (size_t)K))) return rc;
    if ((rc = ensure(e, e->d_centred, e->centred_cap, (size_t)M * K * 3))) return rc;
    if ((rc = ensure(e, e->d_G, e->g_cap, (size_t)e->m_pad))) return rc;
    e->use_i8 = e->use_gram && (e->gram_variant == 10 || e->gram_variant == -1);
    if (e->use_gram && !e->use_i8)
        if ((rc = ensure(e, e->d_packed, e->packed_cap, (size_t)e->m_pad * k_pad * 3))) return rc;
    if (e->use_i8) {
        if (ct::oz_kernel_mode(K)) {
            if ((rc = ensure(e, e->d_packT, e->packT_cap, ct::oz_packA_bytes((int)M, K)))) return rc;
        } else {
            if ((rc = ensure(e, e->d_packA, e->packA_cap, ct::oz_packA_bytes((int)M, K)))) return rc;
        }
        if ((rc = ensure(e, e->d_packB, e->packB_cap, ct::oz_packB_bytes((int)M, K)))) return rc;
        if ((rc = ensure(e, e->d_Gq, e->gq_cap, ct::oz_g_entries((int)M)))) return rc;
    }

    CT_CUDA(e, cudaEventRecord(e->ev_a, e->s_compute));
    CT_CUDA(e, cudaMemcpyAsync(e->d_raw, coords, sizeof(double) * (size_t)M * N * 3, cudaMemcpyHostToDevice, e->s_compute));
    CT_CUDA(e, cudaMemcpyAsync(e->d_keep, keep.data(), sizeof(int) * (size_t)K, cudaMemcpyHostToDevice, e->s_compute));
    CT_CUDA(e, cudaEventRecord(e->ev_b, e->s_compute));
    CT_CUDA(e, ct::launch_pack(e->d_raw, e->d_keep, (int)M, N, K, ncen, e->m_pad, e->kq_total, e->d_centred, e->d_G,
                               (e->use_gram && !e->use_i8) ? e->d_packed : nullptr, e->sm_count, e->s_compute));
    if (e->use_i8)
        CT_CUDA(e, ct::launch_quant_pack(e->d_centred, (int)M, K, e->d_packA, e->d_packB, e->d_packT, e->d_Gq, e->d_qinfo,
                                         e->d_maxbits, e->sm_count, e->s_compute));
    CT_CUDA(e, cudaEventRecord(e->ev_c, e->s_compute));
    CT_CUDA(e, cudaStreamSynchronize(e->s_compute));  // `keep` (host vector) must outlive the copy
    if (e->use_i8) {
        double qi[4];
        CT_CUDA(e, cudaMemcpy(qi, e->d_qinfo, sizeof qi, cudaMemcpyDeviceToHost));
        e->quant_f = qi[1];
        if (e->gram_variant == 10 && e->quant_f < 0)
            return fail(e, CT_ERR_UNSUPPORTED, "the int8 Gram kernel was forced but the coordinates are not finite or span "
                                               "more than 2^30 A");
        if (e->gram_variant == -1 && e->quant_f < e->min_quant_f) {
            // coordinates too far from the centroid for 31-bit fixed point at the accuracy bound: FP64 tensor path
            e->use_i8 = false;
            if ((rc = ensure(e, e->d_packed, e->packed_cap, (size_t)e->m_pad * k_pad * 3))) return rc;
            CT_CUDA(e, ct::launch_pack(e->d_raw, e->d_keep, (int)M, N, K, ncen, e->m_pad, e->kq_total, e->d_centred, e->d_G,
                                       e->d_packed, e->sm_count, e->s_compute));
            CT_CUDA(e, cudaStreamSynchronize(e->s_compute));
        }
    }
    e->stats.gram_path = e->use_gram ? (e->use_i8 ? 2 : 1) : 0;
    e->stats.quant_step = e->use_i8 ? ldexp(1.0, -(int)e->quant_f) : 0.0;
    float ms = 0.f;
    CT_CUDA(e, cudaEventElapsedTime(&ms, e->ev_a, e->ev_b)); e->stats.h2d_ms = ms;
    CT_CUDA(e, cudaEventElapsedTime(&ms, e->ev_b, e->ev_c)); e->stats.pack_ms = ms;
    e->stats.kept_atoms = K;

    if (e->plan) { ct::pair_plan_free(e->plan); e->plan = nullptr; }
    {
        std::string perr;
        rc = ct::pair_plan_build(&e->plan, kept_z.data(), K, natoms, &e->opts, perr);
        if (rc) return fail(e, rc, perr);
    }
    e->loaded = true;
    return CT_OK;
}

int ct_pack_resident(ct_engine* e, ct_stats* stats) {
    if (!e) return fail(nullptr, CT_ERR_ARG, "engine is NULL");
    if (!e->loaded) return fail(e, CT_ERR_STATE, "no frames loaded (call ct_load_frames first)");
    CT_CUDA(e, cudaSetDevice(e->device));
    const int ncen = e->opts.solute_natoms ? e->natoms : e->K;
    CT_CUDA(e, cudaEventRecord(e->ev_a, e->s_compute));
    CT_CUDA(e, ct::launch_pack(e->d_raw, e->d_keep, (int)e->M, e->N, e->K, ncen, e->m_pad, e->kq_total, e->d_centred, e->d_G,
                               (e->use_gram && !e->use_i8) ? e->d_packed : nullptr, e->sm_count, e->s_compute));
    if (e->use_i8)
        CT_CUDA(e, ct::launch_quant_pack(e->d_centred, (int)e->M, e->K, e->d_packA, e->d_packB, e->d_packT, e->d_Gq, e->d_qinfo,
                                         e->d_maxbits, e->sm_count, e->s_compute));
    CT_CUDA(e, cudaEventRecord(e->ev_b, e->s_compute));
    CT_CUDA(e, cudaStreamSynchronize(e->s_compute));
    float ms = 0.f;
    CT_CUDA(e, cudaEventElapsedTime(&ms, e->ev_a, e->ev_b));
    e->stats.pack_ms = ms;
    if (stats) *stats = e->stats;
    return CT_OK;
}

// One chunk of rows into `dst` (device), asynchronously on s_compute.
static int compute_rows_async(ct_engine* e, int row_begin, int row_end, double* dst, long long idx_base,
                              long long n_entries, int64_t* launches, cudaEvent_t after_main = nullptr) {
    if (e->use_i8) {
        CT_CUDA(e, cudaMemsetAsync(e->d_fix_count, 0, sizeof(unsigned int), e->s_compute));
        ct::OzParams p{};
        p.packA = e->d_packA; p.packB = e->d_packB; p.packT = e->d_packT; p.Gq = e->d_Gq; p.qinfo = e->d_qinfo;
        p.out = dst; p.idx_base = idx_base; p.M = (int)e->M;
        p.row_begin = row_begin; p.row_end = row_end;
        p.inv_k2 = 2.0 / (double)e->K; p.flag_rel = e->flag_rel;
        p.fix.pairs = e->d_fix_pairs; p.fix.count = e->d_fix_count; p.fix.capacity = e->fix_cap;
        CT_CUDA(e, ct::launch_gram_i8(p, e->K, e->sm_count, e->s_compute));
        if (after_main) CT_CUDA(e, cudaEventRecord(after_main, e->s_compute));
        CT_CUDA(e, ct::launch_fixup(e->d_centred, (int)e->M, e->K, p.fix, dst, idx_base, n_entries, e->d_counters,
                                    e->sm_count, e->s_compute));
        *launches += 2;
    } else if (e->use_gram) {
        CT_CUDA(e, cudaMemsetAsync(e->d_fix_count, 0, sizeof(unsigned int), e->s_compute));
        ct::GramParams p;
        p.packed = e->d_packed; p.G = e->d_G; p.out = dst; p.idx_base = idx_base; p.M = (int)e->M;
        p.kq_total = e->kq_total;
        p.row_begin = row_begin; p.row_end = row_end;
        p.inv_k2 = 2.0 / (double)e->K; p.flag_rel = e->flag_rel;
        p.fix.pairs = e->d_fix_pairs; p.fix.count = e->d_fix_count; p.fix.capacity = e->fix_cap;
        CT_CUDA(e, ct::launch_gram(p, e->gram_variant < 0 ? -1 : e->gram_variant, e->sm_count, e->s_compute));
        if (after_main) CT_CUDA(e, cudaEventRecord(after_main, e->s_compute));
        CT_CUDA(e, ct::launch_fixup(e->d_centred, (int)e->M, e->K, p.fix, dst, idx_base, n_entries, e->d_counters,
                                    e->sm_count, e->s_compute));
        *launches += 2;
    } else {
        CT_CUDA(e, ct::launch_pairs_rows(e->plan, e->d_centred, (int)e->M, e->K, row_begin, row_end, dst, idx_base,
                                         e->d_counters, e->sm_count, e->s_compute));
        if (after_main) CT_CUDA(e, cudaEventRecord(after_main, e->s_compute));
        *launches += 1;
    }
    return CT_OK;
}

static int check_rows(ct_engine* e, int64_t& row_begin, int64_t& row_end) {
    if (!e) return fail(nullptr, CT_ERR_ARG, "engine is NULL");
    if (!e->loaded) return fail(e, CT_ERR_STATE, "no frames loaded (call ct_load_frames first)");
    if (row_begin < 0 || row_end > e->M || row_begin > row_end) return fail(e, CT_ERR_ARG, "bad row range");
    return CT_OK;
}

static int finish_stats(ct_engine* e, int64_t pairs, int64_t launches) {
    unsigned long long c[4];
    CT_CUDA(e, cudaMemcpy(c, e->d_counters, sizeof c, cudaMemcpyDeviceToHost));
    e->stats.pairs = pairs;
    e->stats.launches = launches;
    e->stats.flagged = (int64_t)c[0];
    e->stats.lsap_steps = (int64_t)c[1];
    if (e->use_i8) {
        // the int8 Gram launches mark their B operand L2-persisting (csrc/ozaki.cu launch_ts); the stream is idle here, so
        // hand the set-aside lines back before the consumers (clustering, matrix reductions) run
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);   // and the set-aside itself: left in place it slows the pack
        cudaGetLastError();                                       // kernels of the next step 2.6x (1.1 -> 2.9 ms at 100k x 300)
    }
    return CT_OK;
}

int ct_rmsd_rows_device(ct_engine* e, int64_t row_begin, int64_t row_end, const double** out_dev, ct_stats* stats) {
    int rc = check_rows(e, row_begin, row_end);
    if (rc) return rc;
    CT_CUDA(e, cudaSetDevice(e->device));
    const int64_t n = ct_rows_pairs(e->M, row_begin, row_end);
    e->slab_full_M = 0;
    e->reserved_M = 0;
    if ((rc = ensure(e, e->d_slab, e->slab_cap, (size_t)std::max<int64_t>(n, 1)))) return rc;
    CT_CUDA(e, cudaMemsetAsync(e->d_counters, 0, sizeof(unsigned long long) * 4, e->s_compute));
    int64_t launches = 0;
    CT_CUDA(e, cudaEventRecord(e->ev_a, e->s_compute));
    CT_CUDA(e, cudaEventRecord(e->ev_c, e->s_compute));
    if (n > 0) {
        rc = compute_rows_async(e, (int)row_begin, (int)row_end, e->d_slab,
                                ct::condensed_row_base(e->M, row_begin), n, &launches, e->ev_c);
        if (rc) return rc;
    }
    CT_CUDA(e, cudaEventRecord(e->ev_b, e->s_compute));
    CT_CUDA(e, cudaStreamSynchronize(e->s_compute));
    float ms = 0.f, ms_main = 0.f;
    CT_CUDA(e, cudaEventElapsedTime(&ms, e->ev_a, e->ev_b));
    CT_CUDA(e, cudaEventElapsedTime(&ms_main, e->ev_a, e->ev_c));
    e->stats.kernel_ms = ms; e->stats.total_ms = ms; e->stats.d2h_ms = 0.0;
    e->stats.gram_ms = e->use_gram ? ms_main : 0.0; e->stats.pair_ms = e->use_gram ? ms - ms_main : ms;
    if ((rc = finish_stats(e, n, launches))) return rc;
    if (row_begin == 0 && row_end >= e->M - 1) e->slab_full_M = e->M;
    if (out_dev) *out_dev = e->d_slab;
    if (stats) *stats = e->stats;
    return CT_OK;
}

int ct_copy_slab(ct_engine* e, int64_t offset, int64_t count, double* out) {
    if (!e || !out) return fail(e, CT_ERR_ARG, "ct_copy_slab: NULL argument");
    if (offset < 0 || count < 0 || (size_t)(offset + count) > e->slab_cap) return fail(e, CT_ERR_ARG, "ct_copy_slab: range");
    CT_CUDA(e, cudaSetDevice(e->device));
    // Pinned destinations take the direct copy.  A pageable destination (a plain numpy array: the one-shot path) would go
    // through the driver's single-threaded staging at ~4.5 GB/s, first-touch page faults included; here 32 MiB chunks land
    // in a pinned ring at PCIe rate and a handful of host threads move each chunk on while the next one is in flight
    // (the page faults are taken in parallel too).
    cudaPointerAttributes attr{};
    const bool pageable = cudaPointerGetAttributes(&attr, out) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
    cudaGetLastError();
    constexpr size_t CHUNK = (size_t)4 << 20;   // doubles: 32 MiB
    if (!pageable || (size_t)count < 2 * CHUNK) {
        CT_CUDA(e, cudaMemcpy(out, e->d_slab + offset, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost));
        return CT_OK;
    }
    for (int k = 0; k < 2; ++k) {
        if (!e->h_stage[k]) CT_CUDA(e, cudaHostAlloc((void**)&e->h_stage[k], CHUNK * sizeof(double), cudaHostAllocDefault));
        if (!e->ev_stage[k]) CT_CUDA(e, cudaEventCreateWithFlags(&e->ev_stage[k], cudaEventDisableTiming));
    }
    CT_CUDA(e, cudaStreamSynchronize(e->s_compute));
    const double* src = e->d_slab + offset;
    const size_t total = (size_t)count, nchunks = (total + CHUNK - 1) / CHUNK;
    int nt = (int)std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : (nt > 16 ? 16 : nt);
    auto issue = [&](size_t c) -> cudaError_t {
        const size_t lo = c * CHUNK, n = std::min(CHUNK, total - lo);
        cudaError_t rc = cudaMemcpyAsync(e->h_stage[c & 1], src + lo, n * sizeof(double), cudaMemcpyDeviceToHost, e->s_copy);
        return rc != cudaSuccess ? rc : cudaEventRecord(e->ev_stage[c & 1], e->s_copy);
    };
    CT_CUDA(e, issue(0));
    for (size_t c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) CT_CUDA(e, issue(c + 1));        // lands in the other buffer while this one is moved on
        CT_CUDA(e, cudaEventSynchronize(e->ev_stage[c & 1]));
        const size_t lo = c * CHUNK, n = std::min(CHUNK, total - lo);
        const double* from = e->h_stage[c & 1];
        double* to = out + lo;
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; ++t)
            pool.emplace_back([=] { memcpy(to + n * t / nt, from + n * t / nt, (n * (t + 1) / nt - n * t / nt) * sizeof(double)); });
        memcpy(to, from, (n / nt) * sizeof(double));
        for (auto& th : pool) th.join();
    }
    return CT_OK;
}

// Rows [row_begin, row_end) in chunks of about chunk_entries matrix entries: chunk c+1 is computed while chunk c leaves
// the two-deep device ring -- to host memory `out` (may be NULL) over PCIe and, when `root` is given, into root's reserved
// whole-matrix buffer at the rows' own place in the condensed vector (in place on root's own GPU, cudaMemcpyPeerAsync over
// NVLink from the others).
static int rows_stream(ct_engine* e, int64_t row_begin, int64_t row_end, double* out, ct_engine* root, ct_stats* stats) {
    int rc = check_rows(e, row_begin, row_end);
    if (rc) return rc;
    const int64_t n = ct_rows_pairs(e->M, row_begin, row_end);
    if (n > 0 && !out && !root) return fail(e, CT_ERR_ARG, "ct_rmsd_rows: out is NULL");
    CT_CUDA(e, cudaSetDevice(e->device));
    double* root_dst = nullptr;
    if (root) {
        if (root->reserved_M != e->M || !root->d_slab)
            return fail(e, CT_ERR_STATE, "ct_rmsd_rows_gather: call ct_matrix_reserve(root, M) with this trajectory's frame count first");
        root_dst = root->d_slab + ct::condensed_row_base(e->M, row_begin);
        if (root->device != e->device) {
            int can = 0;
            CT_CUDA(e, cudaDeviceCanAccessPeer(&can, e->device, root->device));
            if (can) {   // direct NVLink path; without it cudaMemcpyPeerAsync still works, staged by the driver
                const cudaError_t pe = cudaDeviceEnablePeerAccess(root->device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CT_CUDA(e, pe);
                cudaGetLastError();
            }
        }
    }
    // Row chunks aligned to the Gram block height so that no block of frame pairs is computed twice.
    const int align = e->use_i8 ? ct::oz_row_align() : (e->use_gram ? ct::gram_row_align(e->gram_variant) : 1);
    std::vector<int64_t> cuts;
    cuts.push_back(row_begin);
    {
        int64_t r = row_begin;
        const int64_t last = std::min<int64_t>(row_end, e->M - 1);
        // the first chunks are small so that the device->host stream starts early, then they double
        int64_t target = std::max<int64_t>((int64_t)e->chunk_entries / 16, 1);
        while (r < last) {
            int64_t acc = 0, r2 = r;
            while (r2 < last && acc < target) {
                int64_t nxt = std::min<int64_t>(last, (r2 / align + 1) * align);
                acc += ct_rows_pairs(e->M, r2, nxt);
                r2 = nxt;
            }
            cuts.push_back(r2);
            r = r2;
            target = std::min<int64_t>(target * 2, (int64_t)e->chunk_entries);
        }
    }
    size_t biggest = 1;
    for (size_t c = 0; c + 1 < cuts.size(); ++c)
        biggest = std::max<size_t>(biggest, (size_t)ct_rows_pairs(e->M, cuts[c], cuts[c + 1]));
    if (biggest > e->chunk_cap || !e->d_chunk[0]) {
        for (int k = 0; k < 2; ++k) {
            if (e->d_chunk[k]) { CT_CUDA(e, cudaFree(e->d_chunk[k])); e->d_chunk[k] = nullptr; }
            CT_CUDA(e, cudaMalloc(&e->d_chunk[k], biggest * sizeof(double)));
        }
        e->chunk_cap = biggest;
    }
    CT_CUDA(e, cudaMemsetAsync(e->d_counters, 0, sizeof(unsigned long long) * 4, e->s_compute));
    int64_t launches = 0;
    CT_CUDA(e, cudaEventRecord(e->ev_a, e->s_compute));
    int64_t written = 0;
    bool used[2] = {false, false};
    const bool trace = getenv("CLUSTTRAJ_B200_TRACE_COPY") != nullptr;   // development: per-chunk copy timeline on stderr
    std::vector<cudaEvent_t> tr_ev;
    std::vector<int64_t> tr_n;
    for (size_t c = 0; c + 1 < cuts.size(); ++c) {
        const int b = (int)(c & 1);
        const int64_t cn = ct_rows_pairs(e->M, cuts[c], cuts[c + 1]);
        if (cn == 0) continue;
        if (used[b]) {   // both consumers of this ring slot must be done with it
            if (out) CT_CUDA(e, cudaStreamWaitEvent(e->s_compute, e->ev_copy_done[b], 0));
            if (root) CT_CUDA(e, cudaStreamWaitEvent(e->s_compute, e->ev_peer_done[b], 0));
        }
        used[b] = true;
        rc = compute_rows_async(e, (int)cuts[c], (int)cuts[c + 1], e->d_chunk[b],
                                ct::condensed_row_base(e->M, cuts[c]), cn, &launches);
        if (rc) return rc;
        CT_CUDA(e, cudaEventRecord(e->ev_chunk_done[b], e->s_compute));
        if (out) {
            CT_CUDA(e, cudaStreamWaitEvent(e->s_copy, e->ev_chunk_done[b], 0));
            if (trace) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, e->s_copy); tr_ev.push_back(ev); }
            CT_CUDA(e, cudaMemcpyAsync(out + written, e->d_chunk[b], sizeof(double) * (size_t)cn, cudaMemcpyDeviceToHost,
                                       e->s_copy));
            if (trace) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, e->s_copy); tr_ev.push_back(ev); tr_n.push_back(cn); }
            CT_CUDA(e, cudaEventRecord(e->ev_copy_done[b], e->s_copy));
        }
        if (root) {
            CT_CUDA(e, cudaStreamWaitEvent(e->s_peer, e->ev_chunk_done[b], 0));
            if (root->device == e->device)
                CT_CUDA(e, cudaMemcpyAsync(root_dst + written, e->d_chunk[b], sizeof(double) * (size_t)cn, cudaMemcpyDeviceToDevice,
                                           e->s_peer));
            else
                CT_CUDA(e, cudaMemcpyPeerAsync(root_dst + written, root->device, e->d_chunk[b], e->device,
                                               sizeof(double) * (size_t)cn, e->s_peer));
            CT_CUDA(e, cudaEventRecord(e->ev_peer_done[b], e->s_peer));
        }
        written += cn;
    }
    CT_CUDA(e, cudaEventRecord(e->ev_b, e->s_compute));
    for (int b = 0; b < 2; ++b) {
        if (!used[b]) continue;
        if (out) CT_CUDA(e, cudaStreamWaitEvent(e->s_compute, e->ev_copy_done[b], 0));
        if (root) CT_CUDA(e, cudaStreamWaitEvent(e->s_compute, e->ev_peer_done[b], 0));
    }
    CT_CUDA(e, cudaEventRecord(e->ev_c, e->s_compute));
    CT_CUDA(e, cudaStreamSynchronize(e->s_compute));
    CT_CUDA(e, cudaStreamSynchronize(e->s_copy));
    CT_CUDA(e, cudaStreamSynchronize(e->s_peer));
    float ms = 0.f;
    if (trace) {
        for (size_t c = 0; c < tr_n.size(); ++c) {
            float t0 = 0.f, t1 = 0.f;
            cudaEventElapsedTime(&t0, e->ev_a, tr_ev[2 * c]);
            cudaEventElapsedTime(&t1, e->ev_a, tr_ev[2 * c + 1]);
            fprintf(stderr, "copy %2zu: %9lld entries  start %8.3f ms  end %8.3f ms  %6.1f GB/s\n", c, (long long)tr_n[c], t0, t1,
                    8e-6 * (double)tr_n[c] / (t1 - t0));
        }
        for (cudaEvent_t ev : tr_ev) cudaEventDestroy(ev);
    }
    CT_CUDA(e, cudaEventElapsedTime(&ms, e->ev_a, e->ev_b));
    e->stats.kernel_ms = ms;  // compute stream busy time (includes waits for buffer reuse)
    e->stats.gram_ms = e->use_gram ? ms : 0.0; e->stats.pair_ms = e->use_gram ? 0.0 : ms;
    CT_CUDA(e, cudaEventElapsedTime(&ms, e->ev_a, e->ev_c));
    e->stats.total_ms = ms;
    e->stats.d2h_ms = ms;
    if ((rc = finish_stats(e, n, launches))) return rc;
    if (stats) *stats = e->stats;
    return CT_OK;
}

int ct_rmsd_rows(ct_engine* e, int64_t row_begin, int64_t row_end, double* out, ct_stats* stats) {
    return rows_stream(e, row_begin, row_end, out, nullptr, stats);
}

int ct_matrix_reserve(ct_engine* e, int64_t M) {
    if (!e) return fail(nullptr, CT_ERR_ARG, "engine is NULL");
    if (M < 2 || M > (int64_t)1 << 30) return fail(e, CT_ERR_ARG, "ct_matrix_reserve: need 2 <= M <= 2^30 frames");
    CT_CUDA(e, cudaSetDevice(e->device));
    e->slab_full_M = 0;
    e->reserved_M = 0;
    int rc = ensure(e, e->d_slab, e->slab_cap, (size_t)(M * (M - 1) / 2));
    if (rc) return rc;
    e->reserved_M = M;
    return CT_OK;
}

int ct_rmsd_rows_gather(ct_engine* e, int64_t row_begin, int64_t row_end, double* out, ct_engine* root, ct_stats* stats) {
    if (!root) return fail(e, CT_ERR_ARG, "ct_rmsd_rows_gather: root is NULL");
    return rows_stream(e, row_begin, row_end, out, root, stats);
}

int ct_matrix_commit(ct_engine* e, int64_t M) {
    if (!e) return fail(nullptr, CT_ERR_ARG, "engine is NULL");
    if (e->reserved_M != M || M < 2) return fail(e, CT_ERR_STATE, "ct_matrix_commit: no matrix of M frames was reserved on this engine");
    CT_CUDA(e, cudaSetDevice(e->device));
    CT_CUDA(e, cudaDeviceSynchronize());   // peer copies were synchronised by their senders; this orders them before our streams
    e->slab_full_M = M;
    return CT_OK;
}

int ct_rmsd_condensed(ct_engine* e, const double* coords, const int32_t* atomic_nums, int64_t M, int32_t N,
                      const ct_opts* opts, double* out, ct_stats* stats) {
    int rc = ct_load_frames(e, coords, atomic_nums, M, N, opts);
    if (rc) return rc;
    return ct_rmsd_rows(e, 0, M, out, stats);
}

static int pair_values_impl(ct_engine* e, const int64_t* pairs, int64_t n_pairs, double* values, int32_t* perms,
                            double* transforms, bool ordered, ct_stats* stats, const char* who) {
    if (!e) return fail(nullptr, CT_ERR_ARG, "engine is NULL");
    if (!e->loaded) return fail(e, CT_ERR_STATE, "no frames loaded (call ct_load_frames first)");
    if (n_pairs < 0 || (n_pairs > 0 && (!pairs || !values))) return fail(e, CT_ERR_ARG, std::string(who) + ": NULL argument");
    for (int64_t t = 0; t < n_pairs; ++t) {
        const int64_t i = pairs[2 * t], j = pairs[2 * t + 1];
        if (i < 0 || j < 0 || i >= e->M || j >= e->M || i == j || (ordered && i > j))
            return fail(e, CT_ERR_ARG, std::string(who) + (ordered ? ": need 0 <= i < j < M" : ": need two different frames in 0..M-1"));
    }
    if (n_pairs == 0) return CT_OK;
    CT_CUDA(e, cudaSetDevice(e->device));
    struct Scratch {   // freed on every return path
        long long* pairs = nullptr; double* vals = nullptr; int* perms = nullptr; double* xf = nullptr;
        ~Scratch() { cudaFree(pairs); cudaFree(vals); cudaFree(perms); cudaFree(xf); }
    } d;
    CT_CUDA(e, cudaMalloc(&d.pairs, sizeof(long long) * 2 * (size_t)n_pairs));
    CT_CUDA(e, cudaMalloc(&d.vals, sizeof(double) * (size_t)n_pairs));
    if (perms) CT_CUDA(e, cudaMalloc(&d.perms, sizeof(int) * (size_t)n_pairs * e->K));
    if (transforms) CT_CUDA(e, cudaMalloc(&d.xf, sizeof(double) * 12 * (size_t)n_pairs));
    CT_CUDA(e, cudaMemcpy(d.pairs, pairs, sizeof(long long) * 2 * (size_t)n_pairs, cudaMemcpyHostToDevice));
    CT_CUDA(e, cudaMemsetAsync(e->d_counters, 0, sizeof(unsigned long long) * 4, e->s_compute));
    CT_CUDA(e, cudaEventRecord(e->ev_a, e->s_compute));
    CT_CUDA(e, ct::launch_pairs_list(e->plan, e->d_centred, (int)e->M, e->K, d.pairs, n_pairs, d.vals, d.perms,
                                     e->d_counters, e->sm_count, e->s_compute, d.xf));
    CT_CUDA(e, cudaEventRecord(e->ev_b, e->s_compute));
    CT_CUDA(e, cudaStreamSynchronize(e->s_compute));
    float ms = 0.f;
    CT_CUDA(e, cudaEventElapsedTime(&ms, e->ev_a, e->ev_b));
    e->stats.kernel_ms = ms; e->stats.pair_ms = ms; e->stats.gram_ms = 0.0; e->stats.total_ms = ms; e->stats.d2h_ms = 0.0;
    CT_CUDA(e, cudaMemcpy(values, d.vals, sizeof(double) * (size_t)n_pairs, cudaMemcpyDeviceToHost));
    if (perms) CT_CUDA(e, cudaMemcpy(perms, d.perms, sizeof(int) * (size_t)n_pairs * e->K, cudaMemcpyDeviceToHost));
    if (transforms) CT_CUDA(e, cudaMemcpy(transforms, d.xf, sizeof(double) * 12 * (size_t)n_pairs, cudaMemcpyDeviceToHost));
    int rc = finish_stats(e, n_pairs, 1);
    if (rc) return rc;
    if (stats) *stats = e->stats;
    return CT_OK;
}

int ct_set_frame_elements(ct_engine* e, const int32_t* elements, int64_t M, int32_t K) {
    if (!e) return fail(nullptr, CT_ERR_ARG, "engine is NULL");
    if (!e->loaded) return fail(e, CT_ERR_STATE, "ct_set_frame_elements: load frames first");
    if (!elements || M != e->M || K != e->K)
        return fail(e, CT_ERR_ARG, "ct_set_frame_elements: elements must be [M][K] for the loaded frames (K = kept atoms)");
    if (!e->plan || e->opts.reorder_mode == 0) return CT_OK;   // the elements only matter to a reorder
    if (e->opts.no_hydrogen)
        return fail(e, CT_ERR_UNSUPPORTED, "ct_set_frame_elements: apply the per-frame hydrogen masks on the host and load the "
                                           "compacted frames with no_hydrogen = 0 (clusttraj_b200.distmat.frame_invariant_view)");
    CT_CUDA(e, cudaSetDevice(e->device));
    std::string perr;
    const int rc = ct::pair_plan_frame_elements(e->plan, elements, (int)M, (int)K, perr);
    if (rc) return fail(e, rc, perr.c_str());
    e->per_frame_elements = true;
    return CT_OK;
}

int ct_pair_values(ct_engine* e, const int64_t* pairs, int64_t n_pairs, double* values, int32_t* perms, ct_stats* stats) {
    return pair_values_impl(e, pairs, n_pairs, values, perms, nullptr, true, stats, "ct_pair_values");
}

int ct_pair_transforms(ct_engine* e, const int64_t* pairs, int64_t n_pairs, double* values, double* transforms,
                       ct_stats* stats) {
    if (n_pairs > 0 && !transforms) return fail(e, CT_ERR_ARG, "ct_pair_transforms: transforms is NULL");
    if (e && e->per_frame_elements)
        return fail(e, CT_ERR_UNSUPPORTED, "ct_pair_transforms: not available for trajectories with per-frame element lists");
    return pair_values_impl(e, pairs, n_pairs, values, nullptr, transforms, false, stats, "ct_pair_transforms");
}

int ct_matrix_consumers(ct_engine* e, const double* distmat, int64_t M, const int32_t* labels, int32_t n_clusters,
                        int64_t* medoids, double* total, ct_stats* stats) {
    if (!e) return fail(nullptr, CT_ERR_ARG, "engine is NULL");
    if (M < 2 || M > (int64_t)1 << 30) return fail(e, CT_ERR_ARG, "ct_matrix_consumers: need 2 <= M <= 2^30 frames");
    if (n_clusters < 0 || (n_clusters > 0 && (!labels || !medoids)))
        return fail(e, CT_ERR_ARG, "ct_matrix_consumers: labels and medoids are required when n_clusters > 0");
    if (!total) return fail(e, CT_ERR_ARG, "ct_matrix_consumers: total is NULL");
    for (int64_t i = 0; n_clusters > 0 && i < M; ++i)
        if (labels[i] < 1 || labels[i] > n_clusters)
            return fail(e, CT_ERR_ARG, "ct_matrix_consumers: cluster labels must lie in 1..n_clusters");
    CT_CUDA(e, cudaSetDevice(e->device));
    const int64_t L = M * (M - 1) / 2;
    int rc;
    CT_CUDA(e, cudaEventRecord(e->ev_a, e->s_compute));
    if (distmat) {
        e->slab_full_M = 0;
        e->reserved_M = 0;
        if ((rc = ensure(e, e->d_slab, e->slab_cap, (size_t)L))) return rc;
        CT_CUDA(e, cudaMemcpyAsync(e->d_slab, distmat, sizeof(double) * (size_t)L, cudaMemcpyHostToDevice, e->s_compute));
        e->slab_full_M = M;
    } else if (e->slab_full_M != M) {
        return fail(e, CT_ERR_STATE, "ct_matrix_consumers: no whole matrix of M frames on the GPU "
                                     "(pass distmat, or call ct_rmsd_rows_device(e, 0, M, ...) first)");
    }
    CT_CUDA(e, cudaEventRecord(e->ev_b, e->s_compute));
    if (n_clusters > 0 && (rc = check_matrix_values(e, L, true, "ct_matrix_consumers"))) return rc;
    if ((rc = ensure(e, e->d_cons, e->cons_cap, (size_t)ct::consumers_scratch_doubles((int)M)))) return rc;
    if (n_clusters > 0) {
        if ((rc = ensure(e, e->d_labels, e->labels_cap, (size_t)M))) return rc;
        if ((rc = ensure(e, e->d_best, e->best_cap, (size_t)n_clusters))) return rc;
        if ((rc = ensure(e, e->d_medoid, e->medoid_cap, (size_t)n_clusters))) return rc;
        CT_CUDA(e, cudaMemcpyAsync(e->d_labels, labels, sizeof(int) * (size_t)M, cudaMemcpyHostToDevice, e->s_compute));
    }
    CT_CUDA(e, ct::launch_consumers(e->d_slab, L, (int)M, e->d_labels, n_clusters, e->d_cons, e->d_best, e->d_medoid,
                                    e->sm_count, e->s_compute));
    CT_CUDA(e, cudaEventRecord(e->ev_c, e->s_compute));
    CT_CUDA(e, cudaMemcpyAsync(total, e->d_cons + ct::consumers_sum_offset((int)M), sizeof(double), cudaMemcpyDeviceToHost,
                               e->s_compute));
    if (n_clusters > 0)
        CT_CUDA(e, cudaMemcpyAsync(medoids, e->d_medoid, sizeof(long long) * (size_t)n_clusters, cudaMemcpyDeviceToHost,
                                   e->s_compute));
    CT_CUDA(e, cudaStreamSynchronize(e->s_compute));
    for (int c = 0; c < n_clusters; ++c)
        if (medoids[c] < 0 || medoids[c] >= M)  // the reference's np.argmin raises on an empty cluster (classify.py:163)
            return fail(e, CT_ERR_ARG, "ct_matrix_consumers: cluster " + std::to_string(c + 1) + " has no members");
    float ms = 0.f;
    CT_CUDA(e, cudaEventElapsedTime(&ms, e->ev_a, e->ev_b)); e->stats.h2d_ms = ms;
    CT_CUDA(e, cudaEventElapsedTime(&ms, e->ev_b, e->ev_c)); e->stats.kernel_ms = ms;
    e->stats.total_ms = e->stats.h2d_ms + e->stats.kernel_ms;
    e->stats.pairs = L; e->stats.launches = n_clusters > 0 ? 5 : 2;
    if (stats) *stats = e->stats;
    return CT_OK;
}

int ct_linkage(ct_engine* e, const double* distmat, int64_t M, int32_t method, double* Z, ct_stats* stats) {
    if (!e) return fail(nullptr, CT_ERR_ARG, "engine is NULL");
    if (M < 2 || M > (int64_t)1 << 30) return fail(e, CT_ERR_ARG, "ct_linkage: need 2 <= M <= 2^30 observations");
    if (!Z) return fail(e, CT_ERR_ARG, "ct_linkage: Z is NULL");
    if (method != 0 && method != 1 && method != 2 && method != 5 && method != 6)
        return fail(e, CT_ERR_UNSUPPORTED, "ct_linkage: methods single (0), complete (1), average (2), ward (5) and weighted (6) "
                                           "run on the GPU; centroid / median stay with SciPy");
    CT_CUDA(e, cudaSetDevice(e->device));
    const int64_t L = M * (M - 1) / 2;
    int rc;
    CT_CUDA(e, cudaEventRecord(e->ev_a, e->s_compute));
    if (distmat) {
        e->slab_full_M = 0;
        e->reserved_M = 0;
        if ((rc = ensure(e, e->d_slab, e->slab_cap, (size_t)L))) return rc;
        CT_CUDA(e, cudaMemcpyAsync(e->d_slab, distmat, sizeof(double) * (size_t)L, cudaMemcpyHostToDevice, e->s_compute));
        e->slab_full_M = M;
    } else if (e->slab_full_M != M) {
        return fail(e, CT_ERR_STATE, "ct_linkage: no whole matrix of M observations on the GPU "
                                     "(pass distmat, or call ct_rmsd_rows_device(e, 0, M, ...) first)");
    }
    CT_CUDA(e, cudaEventRecord(e->ev_b, e->s_compute));
    if ((rc = check_matrix_values(e, L, false, "ct_linkage"))) return rc;
    // scratch for the matrix the algorithm updates (the matrix itself stays intact, like SciPy's input): the square form
    // (M x M, persistent cluster kernel) when it fits, else a condensed copy (step-launch variant)
    {
        size_t want = (size_t)M * (size_t)M;
        size_t free_b = 0, total_b = 0;
        CT_CUDA(e, cudaMemGetInfo(&free_b, &total_b));
        const size_t have = e->d_link ? e->link_cap * sizeof(double) : 0;
        if (want * sizeof(double) > free_b + have) want = (size_t)L;
        if ((rc = ensure(e, e->d_link, e->link_cap, want))) return rc;
    }
    if ((rc = ensure(e, e->d_link_scratch, e->link_scratch_cap, ct::linkage_scratch_bytes((int)M, e->sm_count)))) return rc;
    if (!e->h_link_flag) CT_CUDA(e, cudaHostAlloc((void**)&e->h_link_flag, sizeof(int), cudaHostAllocDefault));
    std::vector<double> zraw(4 * (size_t)(M - 1));
    long long steps = 0;
    CT_CUDA(e, ct::run_linkage(e->d_slab, e->d_link, e->link_cap, (int)M, method, e->d_link_scratch, e->sm_count, zraw.data(),
                               e->h_link_flag, &steps, e->s_compute));
    CT_CUDA(e, cudaEventRecord(e->ev_c, e->s_compute));
    CT_CUDA(e, cudaEventSynchronize(e->ev_c));
    ct::linkage_finish(zraw.data(), (int)M, Z);
    float ms = 0.f;
    CT_CUDA(e, cudaEventElapsedTime(&ms, e->ev_a, e->ev_b)); e->stats.h2d_ms = ms;
    CT_CUDA(e, cudaEventElapsedTime(&ms, e->ev_b, e->ev_c)); e->stats.kernel_ms = ms;
    e->stats.total_ms = e->stats.h2d_ms + e->stats.kernel_ms;
    e->stats.pairs = L; e->stats.launches = steps + 1;
    if (stats) *stats = e->stats;
    return CT_OK;
}

int ct_device_memory(ct_engine* e, int64_t* free_bytes, int64_t* total_bytes) {
    if (!e || !free_bytes || !total_bytes) return fail(e, CT_ERR_ARG, "ct_device_memory: NULL argument");
    CT_CUDA(e, cudaSetDevice(e->device));
    size_t f = 0, t = 0;
    CT_CUDA(e, cudaMemGetInfo(&f, &t));
    *free_bytes = (int64_t)f;
    *total_bytes = (int64_t)t;
    return CT_OK;
}

int ct_get_stats(const ct_engine* e, ct_stats* stats) {
    if (!e || !stats) return CT_ERR_ARG;
    *stats = e->stats;
    return CT_OK;
}

}  // extern "C"
