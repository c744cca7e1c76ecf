/* cfx_b200.h -- C ABI of the B200 charge-flux Ewald electrostatics library (libcfx_b200.so).
 *
 * Boundary being replaced (reference file:line, all under /root/reference):
 *   - CalcCoulForceKernel::initialize(const System&, const CoulForce&)   openmmapi/include/CoulKernels.h:29
 *       -> cfx_create(): takes exactly the parameter arrays CoulForce stores
 *          (openmmapi/include/CoulForce.h:138-149) plus the System's default box
 *          (platforms/reference/src/ReferenceCoulKernels.cpp:400).
 *   - CalcCoulForceKernel::execute(ContextImpl&, bool includeForces, bool includeEnergy)
 *                                                                         openmmapi/include/CoulKernels.h:37
 *       -> cfx_execute() (host buffers, what a Reference/CPU-style platform holds,
 *          platforms/reference/src/ReferenceCoulKernels.cpp:14-27) and cfx_execute_device()
 *          (device buffers, what a CUDA-style platform holds, platforms/cuda/src/CudaCoulKernels.cpp:529-557).
 *   - ~KernelImpl -> cfx_destroy().
 *   - C++ exceptions (OpenMMException) -> int status + cfx_last_error(); the C++ plugin adapter
 *     (openmm_chargeflux_b200/plugin/) rethrows them as OpenMMException.
 *
 * Units follow OpenMM: nm, elementary charge, kJ/mol, radians. All arrays are caller-owned and only
 * touched during the call (the one opt-in exception is CFX_OPT_PIN_CALLER_BUFFERS below); the handle owns
 * every device allocation, stream and CUDA graph.
 * One handle per host thread; no global state besides the per-thread last-error string.
 *
 * There is no CPU fallback: every entry point that computes fails with CFX_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef CFX_B200_H_
#define CFX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFX_OK            0
#define CFX_ERR_ARGUMENT  1   /* bad sizes / indices / null pointers / unsupported box            */
#define CFX_ERR_CUDA      2   /* CUDA runtime error or no usable device                          */
#define CFX_ERR_STATE     3   /* getter called before the data it returns was computed           */

/* Coulomb constant shared by oracle, reference build and kernels. */
#define CFX_ONE_4PI_EPS0 138.935456

/* Index into the energy[] output of cfx_execute*. total = self + recip + direct + excl (PBC,
 * ReferenceCoulKernels.cpp:633) or the all-pairs sum (non-PBC, :436-491; only TOTAL/DIRECT/EXCL used). */
enum { CFX_E_SELF = 0, CFX_E_RECIP = 1, CFX_E_DIRECT = 2, CFX_E_EXCL = 3, CFX_E_TOTAL = 4, CFX_E_COUNT = 5 };

/* Parameter block: one field per CoulForce storage vector (CoulForce.h:138-149), in the layout the
 * getters return them (CoulForce.cpp:28-32,65-68,85-90,104-110,127-136). */
typedef struct cfx_system_desc {
    int32_t        num_particles;
    const double*  charge;              /* [N]    base charge q0                                    */
    const double*  sigma;               /* [N]    LJ sigma (nm), combined as (s_i+s_j)/2             */
    const double*  epsilon;             /* [N]    LJ epsilon (kJ/mol), combined as 4*sqrt(e_i*e_j)   */
    int32_t        num_exceptions;
    const int32_t* exception_pairs;     /* [2X]   excluded pairs (p1,p2)                            */
    int32_t        num_flux_bonds;
    const int32_t* flux_bond_idx;       /* [2nb]  (p1,p2)                                           */
    const double*  flux_bond_params;    /* [2nb]  (k [e/nm], b [nm])                                */
    int32_t        num_flux_angles;
    const int32_t* flux_angle_idx;      /* [3na]  (p1,p2,p3), p2 is the apex                        */
    const double*  flux_angle_params;   /* [2na]  (k [e/rad], theta0 [rad])                         */
    int32_t        num_flux_waters;
    const int32_t* flux_water_idx;      /* [3nw]  (O,H1,H2)                                         */
    const double*  flux_water_params;   /* [5nw]  (k1,k2,kub,b0,ub0)                                */
    double         cutoff;              /* nm; PBC only                                             */
    double         ewald_tol;           /* PBC only                                                 */
    int32_t        use_pbc;
    double         default_box[9];      /* row-major a,b,c; alpha/kmax are fixed from this at create */
} cfx_system_desc;

/* cfx_options.flags. PIN_CALLER_BUFFERS: the caller promises that a positions / forces array it passes to
 * cfx_execute stays allocated until it passes a different array or destroys the handle (true for a platform's own
 * persistent vectors, platforms/reference/src/ReferenceCoulKernels.cpp:14-27). The library then page-locks such an
 * array in place (cudaHostRegister) the second time it sees it, DMA-reads positions from it and accumulates forces
 * into it directly, and unregisters it when a different array arrives or at cfx_destroy. Without the flag (default)
 * every call stages through the handle's own pinned buffers and caller memory is only touched during the call. */
#define CFX_OPT_PIN_CALLER_BUFFERS 1
/* SKIP_DISCARDED_ENERGY: with include_energy == 0 the reference still returns self + direct + exclusion energy, a value
 * OpenMM discards; by default it is reproduced (FP32 pair terms). With this flag such calls
 * return 0 for the direct component and skip its arithmetic. */
#define CFX_OPT_SKIP_DISCARDED_ENERGY 2
/* KMAX_FOLLOWS_BOX: the reference fixes kmax from the System's DEFAULT box at initialize and never revisits it
 * (ReferenceCoulKernels.cpp:403-420), so under a barostat the reciprocal-space error drifts with the box (SURVEY.md
 * section 8 f4). With this flag the same estimator is re-applied to the box of the call whenever the box has changed, and the
 * reciprocal-space plan is rebuilt when kmax moves. Default off = the reference's behaviour. */
#define CFX_OPT_KMAX_FOLLOWS_BOX 4

/* Execution options (all optional; pass NULL for defaults). */
typedef struct cfx_options {
    int32_t device;        /* CUDA device ordinal; -1 = current device                              */
    int32_t shard_rank;    /* this handle's rank in a k-vector / spatial-tile sharded evaluation     */
    int32_t shard_count;   /* number of ranks sharing one evaluation (1 = whole evaluation here)     */
    int32_t use_graph;     /* 1 = replay the step as one CUDA graph (default), 0 = plain launches    */
    int32_t flags;         /* CFX_OPT_* bits                                                         */
    int32_t list_skin_pm;  /* skin of the direct-space candidate lists in picometres: the lists are built for
                              cutoff + skin and reused until an atom has moved skin/2 (the in-cutoff test itself is
                              exact every call). 0 = default (100 pm), negative = rebuild at every evaluation        */
    int32_t reserved[2];
} cfx_options;

typedef struct cfx_handle cfx_handle;

/* Derived Ewald parameters (ReferenceCoulKernels.cpp:401-420). */
typedef struct cfx_ewald_params {
    double  alpha;
    int32_t kmax[3];
    int64_t num_kvectors;   /* half-space count the reference loops over (:519-555)                 */
} cfx_ewald_params;

/* Per-evaluation counters, for benchmarks. */
typedef struct cfx_stats {
    int64_t pairs_in_cutoff;    /* in-cutoff, non-excluded i<j pairs of the last evaluation         */
    int64_t pair_candidates;    /* distance tests the pair kernel executed                          */
    int64_t kernel_launches;    /* kernels of this library launched by the last evaluation (some return at once when the
                                   candidate lists are reused)                                       */
    int32_t cells[3];
    int32_t longest_pair_list;  /* longest per-cluster candidate list built so far (capacity diagnostics) */
    int64_t pair_list_builds;   /* how many evaluations of this handle re-sorted the atoms and rebuilt the lists */
} cfx_stats;

const char* cfx_last_error(void);
int  cfx_device_count(void);

int  cfx_create(const cfx_system_desc* desc, const cfx_options* opts, cfx_handle** out);
void cfx_destroy(cfx_handle* h);
/* New parameter values (charges, sigma, epsilon, flux k/b/theta0/...) for an existing handle with the same particle
 * and flux-term counts and the same index lists; the exception list, cutoff, tolerance and box rule are not re-read.
 * What an `updateParametersInContext` of the API layer would call (the reference has none: SURVEY.md section 8 f4). */
int  cfx_update_parameters(cfx_handle* h, const cfx_system_desc* desc);

/* Host-buffer evaluation (the reference-facing call). positions: [3N] double, unwrapped, nm.
 * box: row-major a,b,c (only read when use_pbc). energy: [CFX_E_COUNT] written. forces: [3N] double,
 * ADDED to (as the reference adds to the platform's force vector, ReferenceCoulKernels.cpp:455-630);
 * may be NULL. Flag semantics mirror the reference, including its quirks:
 * PBC: recip energy only if include_energy, the other components always; pair/recip forces and
 * dE/dq only if include_forces, self dE/dq and the chain rule always. */
int  cfx_execute(cfx_handle* h, const double* positions, const double* box,
                 int include_forces, int include_energy, double* energy, double* forces);

/* Device-buffer evaluation on `stream` (a cudaStream_t passed as void*). d_positions: [3N] double on
 * the handle's device. d_force_fixed: int64 [3][Npad] fixed-point (value*2^32, the OpenMM CUDA
 * platform convention, PBCForce.cu:336-338), ADDED to atomically; Npad = cfx_padded_num_particles().
 * d_dedq_fixed (may be NULL): int64 [Npad] fixed-point dE/dq, ADDED to. d_energy (may be NULL):
 * double[CFX_E_COUNT], ADDED to. Asynchronous: returns after enqueueing. With shard_count > 1 the
 * outputs are this rank's partial sums; the caller all-reduces them (sum) across ranks. */
int  cfx_execute_device(cfx_handle* h, const double* d_positions, const double* box,
                        int include_forces, int include_energy,
                        long long* d_force_fixed, long long* d_dedq_fixed, double* d_energy, void* stream);

/* Evaluation inside a CUDA-platform host (replaces platforms/cuda/src/CudaCoulKernels.cpp:523-660 and
 * kernels/PBCForce.cu:817-825). d_posq: the platform's `real4 posq[padded_num_atoms]` (float4, or double4 when
 * posq_is_double) in the PLATFORM's atom order; d_posq_correction: NULL, or the float4 array of low-order position bits
 * a mixed-precision platform keeps beside a float4 posq; d_atom_index[slot] = user particle index
 * (CudaContext::getAtomIndexArray). d_force_buffers: the platform's 64-bit fixed-point force buffer [3][padded_num_atoms]
 * (value*2^32, PBCForce.cu:336-338), platform order, ADDED to. d_energy_buffer: NULL, or the platform's energy buffer
 * (double when energy_is_double, else float): the total is ADDED to its element 0. Asynchronous on `stream`; the gather
 * into user order, the evaluation and the scatter are replayed as one CUDA graph. */
int  cfx_execute_platform(cfx_handle* h, const void* d_posq, int posq_is_double, const void* d_posq_correction,
                          const int32_t* d_atom_index, int32_t padded_num_atoms, const double* box,
                          int include_forces, int include_energy, unsigned long long* d_force_buffers,
                          void* d_energy_buffer, int energy_is_double, void* stream);

/* The same evaluation for a sharded (multi-GPU) step: d_reduce is this rank's reduction buffer, int64
 * [3*Npad + 8] = fixed-point forces (value*2^32) followed by the CFX_E_* energies as value*2^24 (3 spare slots).
 * The buffer is ZEROED and filled inside the call's CUDA graph, so one step of a sharded evaluation is this call
 * plus one sum all-reduce of d_reduce (NCCL over NVLink); integer sums make the result independent of the order. */
int  cfx_execute_shard(cfx_handle* h, const double* d_positions, const double* box,
                       int include_forces, int include_energy, long long* d_reduce, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU. The reference has none (its CUDA platform binds contexts[0] only,
 * platforms/cuda/src/CudaCoulKernelFactory.cpp:40). An evaluation is split over the ranks of an NCCL communicator
 * -- each rank owns a block of k-vector rows and a slab of direct-space i-clusters -- and completed by one sum
 * all-reduce of the int64 fixed-point reduction buffer over NVLink (integer sums: any reduction order gives the same bits).
 * NCCL is dlopen'ed (libnccl.so.2) on first use.
 *
 * One process (or thread) per GPU: every rank creates its handle with (shard_rank, shard_count), rank 0 draws an id
 * with cfx_comm_get_unique_id and the launcher (MPI, torchrun, ...) hands it to every rank's cfx_comm_init, a collective
 * call. After that cfx_execute() works on the sharded handle -- every rank passes the same positions and receives the
 * whole energy and forces -- and cfx_execute_sharded() is cfx_execute_shard() followed by the in-place all-reduce of
 * d_reduce on `stream`, both replayed as one CUDA graph. */
#define CFX_COMM_ID_BYTES 128
int  cfx_comm_get_unique_id(void* id /*[CFX_COMM_ID_BYTES]*/);
int  cfx_comm_init(cfx_handle* h, const void* id /*[CFX_COMM_ID_BYTES]*/);
int  cfx_comm_size(const cfx_handle* h);                  /* ranks of the handle's communicator, 0 = none */
int  cfx_execute_sharded(cfx_handle* h, const double* d_positions, const double* box,
                         int include_forces, int include_energy, long long* d_reduce, void* stream);

/* One process, several GPUs -- what a plugin inside a single OpenMM process can use (the reference's stub of it:
 * CudaCoulKernels.cpp:477-481). The multi-handle owns one sharded handle and one communicator rank per device;
 * cfx_multi_execute has the contract of cfx_execute. cfx_multi_handle(m, i) exposes rank i's handle to the getters. */
typedef struct cfx_multi cfx_multi;
int  cfx_multi_create(const cfx_system_desc* desc, const int32_t* devices, int32_t num_devices, cfx_multi** out);
void cfx_multi_destroy(cfx_multi* m);
int  cfx_multi_num_devices(const cfx_multi* m);
cfx_handle* cfx_multi_handle(cfx_multi* m, int32_t index);
int  cfx_multi_execute(cfx_multi* m, const double* positions, const double* box,
                       int include_forces, int include_energy, double* energy, double* forces);

int  cfx_padded_num_particles(const cfx_handle* h);
int  cfx_get_ewald_params(const cfx_handle* h, cfx_ewald_params* out);
int  cfx_get_stats(const cfx_handle* h, cfx_stats* out);

/* Parity / debug getters for the last cfx_execute (host arrays, user particle order). */
int  cfx_get_charges(cfx_handle* h, double* q /*[N]*/);
int  cfx_get_dedq(cfx_handle* h, double* dedq /*[N]*/);
int  cfx_num_jacobian_rows(const cfx_handle* h);                      /* P = 4nb + 9na + 9nw           */
int  cfx_get_jacobian(cfx_handle* h, int32_t* dq_idx /*[P]*/, int32_t* dx_idx /*[P]*/, double* val /*[3P]*/);
/* In-cutoff non-excluded pairs (i<j) of the last PBC evaluation, sorted lexicographically.
 * Call with pairs == NULL to get the count. */
int  cfx_get_neighbor_pairs(cfx_handle* h, int32_t* pairs /*[2*count]*/, int64_t capacity, int64_t* count);
/* Exclusion lists as CSR (symmetric, sorted): row_ptr [N+1], cols [2X']. */
int  cfx_get_exclusions(cfx_handle* h, int32_t* row_ptr, int32_t* cols, int64_t capacity, int64_t* count);

/* Timing helpers for bench.py: average device time (ms) of `iters` back-to-back evaluations on
 * device-resident positions (CUDA events on the handle's stream), and a per-kernel breakdown
 * (names joined by ';', times in ms, averaged over iters). */
int  cfx_time_device(cfx_handle* h, const double* d_positions, const double* box,
                     int include_forces, int include_energy, int iters, float* ms_per_eval);
int  cfx_time_kernels(cfx_handle* h, const double* d_positions, const double* box,
                      int include_forces, int include_energy, int iters, char* names, int names_capacity, float* ms, int ms_capacity, int* count);
/* Sustained FP32 FMA throughput of the device (TFLOP/s), the roofline denominator the path is
 * bounded by (MEASURED_PEAKS.json has no FP32 CUDA-core figure). */
int  cfx_measure_fp32_peak(int device, int iters, double* tflops, double* sm_clock_mhz_est);
/* Sustained dense TF32 tensor-core throughput (TFLOP/s) of tcgen05.mma kind::tf32 128x128x8, the roofline denominator
 * of the tensor-core gather (MEASURED_PEAKS.json holds a bf16 figure only). */
int  cfx_measure_tf32_peak(int device, int iters, double* tflops);
/* Sustained dense INT8 tensor-core throughput (TOP/s) of tcgen05.mma kind::i8 128x256x32, the roofline denominator of the
 * integer structure-factor kernel. */
int  cfx_measure_i8_peak(int device, int iters, double* tops);

/* ------------------------------------------------------------------------------------------------
 * MD harness: what sits either side of the path in an OpenMM simulation of
 * flexible molecules -- harmonic bond/angle forces and a velocity-Verlet integrator -- so that the
 * "NVE MD" benchmark configuration can be run and energy conservation checked without OpenMM. The whole
 * step (bonded forces + charge-flux Ewald forces + integration) is device-resident and replayed as one
 * CUDA graph. Units: amu, nm, ps, kJ/mol. E_bond = k/2 (r-r0)^2, E_angle = k/2 (theta-theta0)^2.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cfx_md cfx_md;
int  cfx_md_create(cfx_handle* h, const double* masses /*[N]*/,
                   int32_t num_bonds, const int32_t* bond_idx /*[2nb]*/, const double* bond_params /*[2nb]: k, r0*/,
                   int32_t num_angles, const int32_t* angle_idx /*[3na]*/, const double* angle_params /*[2na]: k, theta0*/,
                   cfx_md** out);
void cfx_md_destroy(cfx_md* md);
int  cfx_md_set_state(cfx_md* md, const double* positions /*[3N]*/, const double* velocities /*[3N] or NULL*/, const double* box);
int  cfx_md_get_state(cfx_md* md, double* positions, double* velocities);
/* steepest descent with a per-atom displacement cap (nm), to relax the synthetic starting structures */
int  cfx_md_minimize(cfx_md* md, int32_t steps, double max_displacement);
/* nsteps of velocity Verlet with time step dt (ps); ms_elapsed (may be NULL) = device time of the steps */
int  cfx_md_step(cfx_md* md, int32_t nsteps, double dt, float* ms_elapsed);
/* energies of the current state: [kinetic, bonded, coulomb+LJ (the CoulForce), total] */
int  cfx_md_energies(cfx_md* md, double* out4);

#ifdef __cplusplus
}
#endif
#endif /* CFX_B200_H_ */
