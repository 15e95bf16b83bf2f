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

template <class G1, class G2, class FR, class FRP>
static int groth16_prove_t(Engine &E, int field, const b200_groth16_pk *pk, const void *d_assignment, size_t num_assign,
                           size_t num_aux, void *d_a, void *d_b, void *d_c, unsigned log_n, void *d_proof,
                           cudaStream_t st) {
    using F1 = typename G1::F;
    using F2 = typename G2::F;
    using M = typename FR::Mem;
    const size_t n = (size_t)1 << log_n;
    const size_t J1 = sizeof(JacobianMem<F1>), J2 = sizeof(JacobianMem<F2>);
    int rc;
    if ((rc = E.g16_h.reserve(n * sizeof(M))) || (rc = E.g16_tmp.reserve(3 * J1 + J2))) return rc;
    char *tmp = E.g16_tmp.as<char>();                 // a_acc | l_acc | h_acc | b_acc
    char *proof = reinterpret_cast<char *>(d_proof);  // A (G1) | B (G2) | C (G1)
    if ((rc = witness_map(E, field, d_a, d_b, d_c, (int)log_n, E.g16_h.p, st))) return rc;
    k_into_repr<FR><<<ceil_div(n, 256), 256, 0, st>>>(E.g16_h.as<M>(), (uint32_t)n);
    LAUNCH_CHECK();
    const char *assign = reinterpret_cast<const char *>(d_assignment);
    const char *aux = assign + (num_assign - num_aux) * sizeof(M);
    const char *aq = reinterpret_cast<const char *>(pk->a_query), *bq = reinterpret_cast<const char *>(pk->b_g2_query);
    // the three G1 MSMs as one pipelined batch (sort / accumulate / tail of consecutive MSMs overlap),
    // then the G2 MSM (same batch when G1 and G2 share the coordinate field, i.e. BW6-761)
    const b200_msm_job g1_jobs[4] = {{aq + sizeof(AffineMem<F1>), assign, num_assign, tmp},
                                     {pk->l_query, aux, num_aux, tmp + J1},
                                     {pk->h_query, E.g16_h.p, n - 1, tmp + 2 * J1},
                                     {bq + sizeof(AffineMem<F2>), assign, num_assign, tmp + 3 * J1}};
    if constexpr (std::is_same<G1, G2>::value) {
        if ((rc = msm_batch<G1>(E, g1_jobs, 4, st, nullptr))) return rc;
    } else {
        if ((rc = msm_batch<G1>(E, g1_jobs, 3, st, nullptr))) return rc;
        if ((rc = msm_native<G2>(E, g1_jobs[3].d_bases_packed, assign, num_assign, tmp + 3 * J1, st))) return rc;
    }
    k_proof_coeff<F1><<<1, 1, 0, st>>>(reinterpret_cast<const AffineMem<F1> *>(aq), reinterpret_cast<const JacobianMem<F1> *>(tmp),
                                       reinterpret_cast<const AffineMem<F1> *>(pk->alpha_g1),
                                       reinterpret_cast<JacobianMem<F1> *>(proof));
    LAUNCH_CHECK();
    k_proof_coeff<F2><<<1, 1, 0, st>>>(reinterpret_cast<const AffineMem<F2> *>(bq),
                                       reinterpret_cast<const JacobianMem<F2> *>(tmp + 3 * J1),
                                       reinterpret_cast<const AffineMem<F2> *>(pk->beta_g2),
                                       reinterpret_cast<JacobianMem<F2> *>(proof + J1));
    LAUNCH_CHECK();
    k_sum_jacobian<F1><<<1, SUM_THREADS, 0, st>>>(reinterpret_cast<const JacobianMem<F1> *>(tmp + J1), 2,
                                        reinterpret_cast<JacobianMem<F1> *>(proof + J1 + J2));
    LAUNCH_CHECK();
    ENGINE_MARK(st);
    return B200_OK;
}

int groth16_prove(Engine &E, int family, const b200_groth16_pk *pk, const void *d_assignment, size_t num_assign,
                  size_t num_aux, void *d_a, void *d_b, void *d_c, unsigned log_n, void *d_proof, cudaStream_t st) {
    if (family == B200_GROTH16_BLS12_377)
        return groth16_prove_t<G1_377, G2_377, Fr253, Fr253Params>(E, B200_FR_BLS12_377, pk, d_assignment, num_assign, num_aux,
                                                                   d_a, d_b, d_c, log_n, d_proof, st);
    return groth16_prove_t<G_761, G_761, Fq377, Fq377Params>(E, B200_FR_BW6_761, pk, d_assignment, num_assign, num_aux, d_a,
                                                             d_b, d_c, log_n, d_proof, st);
}

}  // namespace b200
