// Shared host-side declarations of the engine (see engine.cu for the C-ABI).
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/b200_bls.h"
#include "../../include/bls_snark_sys_compat.h"
#include "msm.cuh"

namespace b200 {

int fail(int code, const char *fmt, ...);
void count_launch();

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) return fail(B200_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));      \
    } while (0)

#define LAUNCH_CHECK()                                                                                   \
    do {                                                                                                 \
        count_launch();                                                                                  \
        cudaError_t e_ = cudaGetLastError();                                                             \
        if (e_ != cudaSuccess) return fail(B200_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e_));  \
    } while (0)

// Stream ordering of the engine-owned, grow-only workspaces: whatever stream the previous asynchronous entry point
// ran on recorded Engine::done; the next user of the workspaces waits for it first -- also when it runs on the
// engine's own stream (a wait for an event recorded on the same stream costs nothing).
#define ENGINE_ORDER(st)                                                                                 \
    do {                                                                                                 \
        if (E.has_pending) CUDA_TRY(cudaStreamWaitEvent((st), E.done, 0));                               \
    } while (0)
#define ENGINE_MARK(st)                                                                                  \
    do {                                                                                                 \
        CUDA_TRY(cudaEventRecord(E.done, (st)));                                                         \
        E.has_pending = true;                                                                            \
    } while (0)

// NVTX range around the host-side issue of a stage (header-only NVTX3: a no-op unless a profiler is attached); the
// three stages of an MSM show up as msm.sort / msm.accumulate / msm.tail on the timeline
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

struct Buffer {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return B200_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4;
        CUDA_TRY(cudaMalloc(&p, want));
        cap = want;
        return B200_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// workspace of one MSM in flight (grow-only); two sets so consecutive MSMs of a batch can overlap
struct MsmWs {
    Buffer counts, offsets, cursor, tile_sums, bins, order, sorted, buckets, partials, window_sums, ones, huge_slices;
    Buffer aff_a, aff_b;             // level outputs of the batched-affine accumulation (ping-pong)
};

// one cached radix-2 domain per scalar field (ntt.cuh): constants, power tables, twiddles omega^i (i < n/2)
struct NttDomain {
    int log_n = -1;
    Buffer consts, pw, tw;
};

struct Engine {
    int device = -1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;      // completion of the last MSM (orders workspace reuse across streams)
    bool has_pending = false;
    bool dead = false;               // torn down by b200_shutdown while another thread still held a reference
    std::mutex mu;
    static constexpr int PIPE_SETS_MAX = 4;
    MsmWs ws[PIPE_SETS_MAX];         // [0], [1]: every path; the others: further sets of the batch pipeline (a tail may lag behind)
    // software pipeline of b200_msm_batch_device: sort / accumulate / tail streams and their hand-over events
    cudaStream_t pipe_stream[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_sorted[PIPE_SETS_MAX] = {}, ev_acc[PIPE_SETS_MAX] = {}, ev_tail[PIPE_SETS_MAX] = {};
    // host-pointer MSM: chunked H2D copies on their own stream, one `ready` event per chunk
    cudaStream_t copy_stream = nullptr;
    static constexpr int MAX_CHUNKS = 16;
    cudaEvent_t ev_chunk[MAX_CHUNKS] = {};
    // chunk-fed MSM in flight (msm_chunks_*): the plan of the whole input and the next chunk's index
    MsmPlan chunk_plan{};
    size_t chunk_index = 0;
    bool chunk_done = false;         // the last chunk has run the tail itself (window groups)
    // window groups of one MSM (msm_split_run): group A's accumulate stream and the hand-over events
    cudaStream_t split_stream = nullptr;
    cudaEvent_t ev_split[3] = {};
    // pageable callers: two pinned staging slots (host threads fill slot k + 1 while slot k crosses PCIe)
    void *pinned[2] = {nullptr, nullptr};
    size_t pinned_cap = 0;
    cudaEvent_t ev_slot[2] = {};
    // staging for the host-pointer API
    Buffer h2d_bases, native_bases, scalars, result;
    // multi-pairing: Miller values, packed G2 staging
    Buffer miller, g2_packed, h2d_g2;
    // *_sharded entry points: per-GPU partial results gathered on the primary GPU
    Buffer gather;
    // NTT domains: [0] BLS12-377 Fr, [1] BW6-761 Fr (= BLS12-377 Fq)
    NttDomain ntt[2];
    // Groth16 prover composite (inst_groth16.cu): quotient coefficients h, partial MSM results
    Buffer g16_h, g16_tmp, g16_part, census;
    // batch-verification composites (inst_verify.cu)
    Buffer v_g1jac, v_g2jac, v_g1aff, v_g2aff, v_sum, v_pairs1, v_pairs2, v_offsets, v_flags;
    // hash-to-G1 (inst_hash.cu): affine multiples of the Bowe-Hopwood generators (built at first use), per-call staging
    Buffer bh_table, hash_ws, sqrt_tables;
    bool bh_ready = false, sqrt_ready = false;
    // optional timing of the dominant kernel (b200_profile_*): event pairs around k_bucket_accumulate
    bool profile = false;
    static constexpr int PROF_SLOTS = 256;
    cudaEvent_t prof_ev[2 * PROF_SLOTS] = {};
    int prof_used = 0;
    uint64_t prof_units = 0;         // scalar-point pairs covered by the recorded launches
};



inline int ceil_div(size_t a, size_t b) { return (int)((a + b - 1) / b); }

// per-curve entry points; each is instantiated in its own translation unit (inst_*.cu)
template <class C> int msm_device(Engine &E, const void *d_bases, size_t stride, const void *d_scalars, size_t n, void *d_out, cudaStream_t st);
template <class C> int msm_native(Engine &E, const void *d_bases, const void *d_scalars, size_t n, void *d_out, cudaStream_t st, size_t n_eff = 0);
template <class C> int scalar_census(Engine &E, const void *d_scalars, size_t n, size_t *n_eff, cudaStream_t st);
template <class C> int msm_batch(Engine &E, const b200_msm_job *jobs, size_t count, cudaStream_t st, const cudaEvent_t *ready = nullptr,
                                 const size_t *n_eff = nullptr);
template <class C> int msm_chunks_begin(Engine &E, size_t n_total, size_t chunk_max, cudaStream_t st);
template <class C> int msm_chunks_add(Engine &E, const void *d_bases_packed, const void *d_scalars, size_t cnt, cudaEvent_t ready, int last,
                                      void *d_out);
template <class C> int msm_chunks_finish(Engine &E, void *d_out, cudaStream_t st);
template <class C> int pack_bases(const void *src_dev, size_t stride, size_t n, void *dst, cudaStream_t st);
template <class C> int sum_jacobian(const void *pts, size_t count, void *out, cudaStream_t st);
template <class C> int sum_jacobian_batch(const void *pts, size_t count, size_t batch, void *out, cudaStream_t st);
template <class C> int fixed_base_mul(Engine &E, const void *base, const void *scalars, size_t n, void *out, cudaStream_t st);
template <class C> int point_runs(Engine &E, const void *base, const void *scalars, size_t runs, size_t run, void *out, cudaStream_t st);
template <class C> int batch_to_affine(const void *jac, size_t n, void *out, cudaStream_t st);
template <class C> int plan_query(size_t n, int *c, int *w, uint32_t *nb);
template <class C> int field_op(int op, const void *a, const void *b, size_t n, void *out, cudaStream_t st);

// BLS12-377 multi-pairing (inst_pairing.cu)
int miller_product(Engine &E, const void *d_g1_packed, const void *d_g2_packed, size_t n, void *d_out, cudaStream_t st);
int final_exp(Engine &E, const void *d_vals, size_t count, void *d_out, int *d_is_one, cudaStream_t st);
int pairing_checks_2(Engine &E, const void *d_g1_packed, const void *d_g2_packed, size_t count, int *d_flags, cudaStream_t st);
// radix-2 NTT / Groth16 witness map (inst_ntt.cu)
int ntt_transform(Engine &E, int field, void *data, int log_n, int inverse, int coset, cudaStream_t st);
int witness_map(Engine &E, int field, void *a, void *b, void *c, int log_n, void *h, cudaStream_t st);
// Groth16 prover arithmetic (inst_groth16.cu)
int groth16_prove(Engine &E, int family, const b200_groth16_pk *pk, const void *d_assignment, size_t num_assign,
                  size_t num_aux, void *d_a, void *d_b, void *d_c, unsigned log_n, void *d_proof, cudaStream_t st);
int groth16_partial(Engine &E, int family, const b200_groth16_pk *pk, const void *d_assignment, size_t num_assign, size_t num_aux,
                    void *d_a, void *d_b, void *d_c, unsigned log_n, unsigned shard, unsigned shards, void *d_partials, cudaStream_t st);
int groth16_assemble(Engine &E, int family, const b200_groth16_pk *pk, const void *d_partials, unsigned shards, void *d_proof,
                     cudaStream_t st);
// BW6-761 pairing / Groth16 verification (inst_bw6_pairing.cu)
int bw6_miller_values(Engine &E, const void *d_g1_packed, const void *d_g2_packed, size_t n, void *d_vals, cudaStream_t st);
int bw6_final_exp(Engine &E, const void *d_vals, size_t count, void *d_out, int *d_is_one, cudaStream_t st);
int bw6_multi_pairing_host(Engine &E, const void *g1, size_t stride1, const void *g2, size_t stride2, size_t n, void *out_fq6,
                           int *out_is_one);
int bw6_groth16_verify(Engine &E, const b200_groth16_vk *vk, const void *proof_a, const void *proof_b, const void *proof_c,
                       const uint64_t *inputs, size_t num_inputs, int *out_verified);
int bw6_groth16_verify_core(Engine &E, char *d_packed, size_t nabc, const void *d_scalars, int *out_verified);
// point decoding and the bls-snark-sys `verify` entry point (inst_epoch_verify.cu)
int decode_points_host(Engine &E, int kind, const void *bytes, size_t n, int subgroup, void *out_packed, int *out_status);
int encode_points_host(Engine &E, int kind, const void *images, size_t n, void *out_bytes);
int epoch_verify(Engine &E, const uint8_t *vk, size_t vk_len, const uint8_t *proof, size_t proof_len, const EpochBlockFFI &first,
                 const EpochBlockFFI &last, int *out_ok, std::string *why);
int epoch_public_inputs(Engine &E, const EpochBlockFFI &first, const EpochBlockFFI &last, std::vector<uint64_t> *inputs, int *ok,
                        std::string *why);
void blake2s_personal(const uint8_t *data, size_t len, const uint8_t personal[8], uint8_t out[32]);
void blake2s_param(const uint8_t *data, size_t len, int digest_len, int fanout, int depth, uint32_t leaf_len, uint64_t node_offset,
                   int inner_len, const uint8_t personal[8], uint8_t *out);
void epoch_block_encode(int cip22, uint16_t index, uint8_t round, const uint8_t *epoch_entropy, const uint8_t *parent_entropy,
                        uint32_t maximum_non_signers, size_t maximum_validators, const uint8_t *keys96, size_t nkeys,
                        std::vector<uint8_t> *inner, std::vector<uint8_t> *extra);
// batched hash-to-G1 (inst_hash.cu)
int hash_to_g1(Engine &E, int hasher, int flags, const uint8_t *domain, size_t domain_len, const b200_hash_input *inputs, size_t n,
               void *out, uint32_t *out_attempts);
// batch-verification flows (inst_verify.cu)
int batch_verify_hashes(Engine &E, const void *signature, const void *pubkeys, const void *hashes, size_t n, int *out_verified);
int batch_verify_strict_many(Engine &E, const b200_strict_batch *batches, size_t count, int *out_verified);
int batch_verify_strict_hash(Engine &E, const void *pubkeys, const void *signatures, const uint64_t *exponents, size_t n,
                             const void *message_hash, int *out_verified);

}  // namespace b200
