// Groth16 prover arithmetic composed on the device: everything ark-groth16 0.1.0
// create_proof_with_reduction_no_zk does AFTER constraint synthesis (SURVEY.md appendix A.4 steps
// 2-4), i.e. what crates/epoch-snark/src/api/prover.rs:78 (outer proof, BW6-761) and :112 (inner
// proof, BLS12-377) spend their arithmetic time on:
//   h       = witness_map(a, b, c)                       (7 radix-2 transforms, ntt.cuh)
//   h_acc   = MSM(h_query, h[0 .. n-1))                  l_acc = MSM(l_query, aux assignment)
//   A       = a_query[0] + MSM(a_query[1..], assignment) + alpha_g1          (r = 0)
//   B       = b_g2_query[0] + MSM(b_g2_query[1..], assignment) + beta_g2     (s = 0)
//   C       = l_acc + h_acc                              (s A + r B1 - r s delta vanish for r = s = 0)
// Constraint synthesis itself (serial symbolic Rust) stays on the host and is out of scope.
#include <type_traits>

#include "engine.cuh"
#include "ntt.cuh"

namespace b200 {

// Montgomery residues -> canonical integers (PrimeField::into_repr), in place
template <class F>
__global__ void __launch_bounds__(256) k_into_repr(typename F::Mem *__restrict__ data, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F one_plain = F::zero();
    one_plain.l[0] = 1u;                             // x R * 1 / R = x
    data[i] = (F::load(ntt_ldg(data + i)) * one_plain).store();
}

// out = q0 + acc + vk  (two packed affine points and one Jacobian), as an arkworks GroupProjective
template <class F>
__global__ void k_proof_coeff(const AffineMem<F> *__restrict__ q0, const JacobianMem<F> *__restrict__ acc,
                              const AffineMem<F> *__restrict__ vk, JacobianMem<F> *__restrict__ out) {
    if (threadIdx.x || blockIdx.x) return;
    Jacobian<F> total = Jacobian<F>::from_ark(*acc);
    launder(total);
    const AffineMem<F> *pts[2] = {q0, vk};
    for (int k = 0; k < 2; k++) {
        Affine<F> a = Affine<F>::from_ark(ldg_mem(pts[k]));
        if (a.is_inf()) continue;
        Jacobian<F> j = {a.x, a.y, F::one()};
        total.add(j);
        launder(total);
    }
    *out = total.to_ark();
}

// contiguous share [lo, hi) of `n` items for shard `k` of `count` (the split of sharded.py's shard_bounds)
static void shard_range(size_t n, unsigned k, unsigned count, size_t *lo, size_t *hi) {
    const size_t base = n / count, rem = n % count;
    *lo = k * base + std::min<size_t>(k, rem);
    *hi = *lo + base + (k < rem ? 1 : 0);
}

// Partial accumulators of one shard: a_acc | l_acc | h_acc | b_acc (3 G1 + 1 G2 GroupProjective images).  The witness
// map runs whole on every shard (it is 7 transforms, the MSMs are the cost that splits); every MSM takes the shard's
// contiguous slice of its (base, scalar) arrays.
template <class G1, class G2, class FR>
static int groth16_partial_t(Engine &E, int field, const b200_groth16_pk *pk, const void *d_assignment, size_t num_assign,
                             size_t num_aux, void *d_a, void *d_b, void *d_c, unsigned log_n, unsigned shard, unsigned shards,
                             void *d_partials, cudaStream_t st) {
    using F1 = typename G1::F;
    using F2 = typename G2::F;
    using M = typename FR::Mem;
    const size_t n = (size_t)1 << log_n;
    const size_t J1 = sizeof(JacobianMem<F1>);
    int rc;
    if ((rc = E.g16_h.reserve(n * sizeof(M)))) return rc;
    char *tmp = reinterpret_cast<char *>(d_partials);
    const char *assign0 = reinterpret_cast<const char *>(d_assignment);
    size_t a_lo, a_hi, l_lo, l_hi, h_lo, h_hi;
    shard_range(num_assign, shard, shards, &a_lo, &a_hi);
    shard_range(num_aux, shard, shards, &l_lo, &l_hi);
    shard_range(n - 1, shard, shards, &h_lo, &h_hi);
    // a witness is mostly bits: count the scalars that actually reach the buckets (neither 0 nor 1) so the a / l / b MSMs
    // get windows sized for them, not for the length of the assignment (the bucket reduce is paid per bucket: at domain 2^22
    // it cost each of these MSMs 17 ms for 50 k dense scalars).  One tiny kernel + a 16-byte read per array, before the
    // witness map is queued so the read waits for nothing.
    size_t eff_a = a_hi - a_lo, eff_l = l_hi - l_lo;
    if ((rc = scalar_census<G1>(E, assign0 + a_lo * sizeof(M), a_hi - a_lo, &eff_a, st)) ||
        (rc = scalar_census<G1>(E, assign0 + (num_assign - num_aux + l_lo) * sizeof(M), l_hi - l_lo, &eff_l, st)))
        return rc;
    if ((rc = witness_map(E, field, d_a, d_b, d_c, (int)log_n, E.g16_h.p, st))) return rc;
    k_into_repr<FR><<<ceil_div(n, 256), 256, 0, st>>>(E.g16_h.as<M>(), (uint32_t)n);
    LAUNCH_CHECK();
    const char *assign = reinterpret_cast<const char *>(d_assignment);
    const char *aux = assign + (num_assign - num_aux) * sizeof(M);
    const char *aq = reinterpret_cast<const char *>(pk->a_query) + sizeof(AffineMem<F1>);
    const char *bq = reinterpret_cast<const char *>(pk->b_g2_query) + sizeof(AffineMem<F2>);
    const char *lq = reinterpret_cast<const char *>(pk->l_query), *hq = reinterpret_cast<const char *>(pk->h_query);
    // the three G1 MSMs as one pipelined batch (sort / accumulate / tail of consecutive MSMs overlap),
    // then the G2 MSM (same batch when G1 and G2 share the coordinate field, i.e. BW6-761)
    const b200_msm_job jobs[4] = {
        {aq + a_lo * sizeof(AffineMem<F1>), assign + a_lo * sizeof(M), a_hi - a_lo, tmp},
        {lq + l_lo * sizeof(AffineMem<F1>), aux + l_lo * sizeof(M), l_hi - l_lo, tmp + J1},
        {hq + h_lo * sizeof(AffineMem<F1>), E.g16_h.as<char>() + h_lo * sizeof(M), h_hi - h_lo, tmp + 2 * J1},
        {bq + a_lo * sizeof(AffineMem<F2>), assign + a_lo * sizeof(M), a_hi - a_lo, tmp + 3 * J1}};
    const size_t n_eff[4] = {eff_a, eff_l, h_hi - h_lo, eff_a};
    if constexpr (std::is_same<G1, G2>::value) {
        if ((rc = msm_batch<G1>(E, jobs, 4, st, nullptr, n_eff))) return rc;
    } else {
        if ((rc = msm_batch<G1>(E, jobs, 3, st, nullptr, n_eff))) return rc;
        if ((rc = msm_native<G2>(E, jobs[3].d_bases_packed, jobs[3].d_scalars, jobs[3].n, tmp + 3 * J1, st, std::max<size_t>(eff_a, 1))))
            return rc;
    }
    ENGINE_MARK(st);
    return B200_OK;
}

// d_partials: `shards` records of (a_acc | l_acc | h_acc | b_acc), e.g. the all-gather of every GPU's share
template <class G1, class G2>
static int groth16_assemble_t(Engine &E, const b200_groth16_pk *pk, const void *d_partials, unsigned shards, void *d_proof,
                              cudaStream_t st) {
    using F1 = typename G1::F;
    using F2 = typename G2::F;
    const size_t J1 = sizeof(JacobianMem<F1>), J2 = sizeof(JacobianMem<F2>), REC = 3 * J1 + J2;
    int rc;
    // regroup: a_0 .. a_{s-1} | (l, h)_0 .. (l, h)_{s-1} | b_0 .. b_{s-1}, then three sums
    if ((rc = E.g16_tmp.reserve((size_t)shards * REC + 2 * J1 + J2))) return rc;
    char *grp = E.g16_tmp.as<char>(), *sums = grp + (size_t)shards * REC;
    const char *src = reinterpret_cast<const char *>(d_partials);
    char *ga = grp, *glh = grp + (size_t)shards * J1, *gb = grp + (size_t)shards * 3 * J1;
    for (unsigned k = 0; k < shards; k++) {
        CUDA_TRY(cudaMemcpyAsync(ga + k * J1, src + k * REC, J1, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(glh + k * 2 * J1, src + k * REC + J1, 2 * J1, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(gb + k * J2, src + k * REC + 3 * J1, J2, cudaMemcpyDeviceToDevice, st));
    }
    char *proof = reinterpret_cast<char *>(d_proof);  // A (G1) | B (G2) | C (G1)
    k_sum_jacobian<F1><<<1, SUM_THREADS, 0, st>>>(reinterpret_cast<const JacobianMem<F1> *>(ga), shards, reinterpret_cast<JacobianMem<F1> *>(sums));
    LAUNCH_CHECK();
    k_sum_jacobian<F2><<<1, SUM_THREADS, 0, st>>>(reinterpret_cast<const JacobianMem<F2> *>(gb), shards, reinterpret_cast<JacobianMem<F2> *>(sums + J1));
    LAUNCH_CHECK();
    k_sum_jacobian<F1><<<1, SUM_THREADS, 0, st>>>(reinterpret_cast<const JacobianMem<F1> *>(glh), 2 * shards,
                                        reinterpret_cast<JacobianMem<F1> *>(proof + J1 + J2));               // C = sum of l and h shares
    LAUNCH_CHECK();
    const char *aq = reinterpret_cast<const char *>(pk->a_query), *bq = reinterpret_cast<const char *>(pk->b_g2_query);
    k_proof_coeff<F1><<<1, 1, 0, st>>>(reinterpret_cast<const AffineMem<F1> *>(aq), reinterpret_cast<const JacobianMem<F1> *>(sums),
                                       reinterpret_cast<const AffineMem<F1> *>(pk->alpha_g1),
                                       reinterpret_cast<JacobianMem<F1> *>(proof));
    LAUNCH_CHECK();
    k_proof_coeff<F2><<<1, 1, 0, st>>>(reinterpret_cast<const AffineMem<F2> *>(bq),
                                       reinterpret_cast<const JacobianMem<F2> *>(sums + J1),
                                       reinterpret_cast<const AffineMem<F2> *>(pk->beta_g2),
                                       reinterpret_cast<JacobianMem<F2> *>(proof + J1));
    LAUNCH_CHECK();
    ENGINE_MARK(st);
    return B200_OK;
}

int groth16_partial(Engine &E, int family, const b200_groth16_pk *pk, const void *d_assignment, size_t num_assign,
                    size_t num_aux, void *d_a, void *d_b, void *d_c, unsigned log_n, unsigned shard, unsigned shards,
                    void *d_partials, cudaStream_t st) {
    if (family == B200_GROTH16_BLS12_377)
        return groth16_partial_t<G1_377, G2_377, Fr253>(E, B200_FR_BLS12_377, pk, d_assignment, num_assign, num_aux, d_a, d_b, d_c,
                                                        log_n, shard, shards, d_partials, st);
    return groth16_partial_t<G_761, G_761, Fq377>(E, B200_FR_BW6_761, pk, d_assignment, num_assign, num_aux, d_a, d_b, d_c, log_n,
                                                  shard, shards, d_partials, st);
}

int groth16_assemble(Engine &E, int family, const b200_groth16_pk *pk, const void *d_partials, unsigned shards, void *d_proof,
                     cudaStream_t st) {
    if (family == B200_GROTH16_BLS12_377) return groth16_assemble_t<G1_377, G2_377>(E, pk, d_partials, shards, d_proof, st);
    return groth16_assemble_t<G_761, G_761>(E, pk, d_partials, shards, d_proof, st);
}

// one GPU: the single shard, assembled in place
int groth16_prove(Engine &E, int family, const b200_groth16_pk *pk, const void *d_assignment, size_t num_assign,
                  size_t num_aux, void *d_a, void *d_b, void *d_c, unsigned log_n, void *d_proof, cudaStream_t st) {
    int rc = E.g16_part.reserve(4 * 288);
    if (rc) return rc;
    if ((rc = groth16_partial(E, family, pk, d_assignment, num_assign, num_aux, d_a, d_b, d_c, log_n, 0, 1, E.g16_part.p, st))) return rc;
    return groth16_assemble(E, family, pk, E.g16_part.p, 1, d_proof, st);
}

}  // namespace b200
