/*
 * rrnn.h -- C ABI of the B200-native RTE+RRTMGP-NN hot path (librrnn_b200.so).
 *
 * Drop-in boundary for the reference's NN gas optics + RTE flux solvers.  Every entry point cites the
 * reference interface it replaces (paths relative to the reference tree).  The reference's own C seam
 * is the bind(C) kernel layer (rte/kernels/mo_rte_solver_kernels.F90:125,546,1530); the type-bound
 * Fortran procedures above it are mirrored by the .F90 files under fortran/ (ISO_C_BINDING veneer), by the C++ mirror rrnn.hpp and by the Python
 * host mirror in rte_rrtmgp_nn_b200/.
 *
 * Conventions
 *  - All real data are fp32 (wp = sp, rte/mo_rte_kind.F90:29-33); integers are 32-bit.
 *  - Array layout is the reference's: g-point fastest, then layer, then column:
 *      Fortran (ngpt,nlay,ncol) == C [ncol][nlay][ngpt];  profiles (nlay,ncol) == C [ncol][nlay].
 *  - Every function returns 0 on success, non-zero on error; rrnn_last_error() returns the message
 *    (thread-local), the counterpart of the reference's character(len=128) error_msg.
 *  - Pointers named *_d are DEVICE pointers valid on the context's device; work is enqueued on the
 *    context's stream and is asynchronous w.r.t. the host.  Entry points ending in _host take HOST
 *    pointers, stage through pinned memory and return after the results are in host memory.
 *  - There is no CPU fallback: every compute entry point fails if no CUDA device is usable.
 */
#ifndef RRNN_H
#define RRNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RRNN_API __attribute__((visibility("default")))

typedef struct rrnn_ctx rrnn_ctx_t;             /* device, stream, persistent workspace                 */
typedef struct rrnn_model rrnn_model_t;         /* rrtmgp_network_type (neural/mod_network_rrtmgp.F90:34-53) */
typedef struct rrnn_kdist rrnn_kdist_t;         /* spectral tables of ty_gas_optics_rrtmgp used on the NN path */
typedef struct rrnn_cloud_lut rrnn_cloud_lut_t; /* ty_cloud_optics LUT state (extensions/cloud_optics/mo_cloud_optics.F90:32-70) */

/* activation codes, neural/mod_layer.F90:64-95 */
enum { RRNN_ACT_LINEAR = 0, RRNN_ACT_SOFTSIGN = 1, RRNN_ACT_RELU = 2, RRNN_ACT_SIGMOID = 3, RRNN_ACT_HARD_SIGMOID = 4 };

/* One gas of ty_gas_concs (rrtmgp/mo_gas_concentrations.F90:50-88): conc is scalar, a per-layer profile
 * or a full (nlay,ncol) field.  `conc` lives in the same memory space as the other arrays of the call
 * (device for *_d entry points, host for *_host); scalars are passed by value in `value`. */
typedef struct {
  char name[32];     /* lower-case gas name, NUL- or blank-padded ("h2o", "o3", "co2", ...) */
  const float* conc; /* ndims 1: [nlay]; ndims 2: [ncol][nlay]; ignored for ndims 0        */
  float value;       /* ndims 0: the volume mixing ratio                                   */
  int ndims;         /* 0, 1 or 2                                                          */
} rrnn_gas_t;

/* ------------------------------------------------------------------------------------------------ */
/* library / context                                                                                */
RRNN_API const char* rrnn_last_error(void);
RRNN_API int rrnn_version(void);
RRNN_API int rrnn_device_count(void);
/* stream: a cudaStream_t; NULL = the legacy default stream (the one PyTorch uses by default). */
RRNN_API int rrnn_ctx_create(int device, void* stream, rrnn_ctx_t** out);
RRNN_API int rrnn_ctx_destroy(rrnn_ctx_t* ctx);
RRNN_API int rrnn_ctx_set_stream(rrnn_ctx_t* ctx, void* stream);
RRNN_API void* rrnn_ctx_stream(rrnn_ctx_t* ctx);
RRNN_API int rrnn_ctx_synchronize(rrnn_ctx_t* ctx);
/* Run-time flags of rte/mo_rte_rrtmgp_config.F90:23-40.  lw_source_bug_compat = 1 (default) reproduces
 * lw_source_noscat ignoring top_at_1 (rte/kernels/mo_rte_solver_kernels.F90:770-773); 0 orients the
 * level sources physically for top_at_1 = false.  solver_wide = 1 (default): the LW no-scattering solver carries four g-points
 * per lane where the shape fits (lw_solver_v7: ngpt >= 128 and a multiple of 4, nlay >= 8); 0: two per lane (lw_solver_v6).
 * The two differ in the order of the sum over g-points only (<= 3e-7 of the flux). */
RRNN_API int rrnn_ctx_set_flag(rrnn_ctx_t* ctx, const char* name, int value);
/* Per-kernel device timing with CUDA events on the context's stream.  rrnn_ctx_profile(ctx, 1) enables and
 * resets; rrnn_ctx_profile_read synchronises and returns the summed duration and launch count of kernel
 * kind 0 = NN gas optics LW, 1 = LW solver, 2 = NN gas optics SW, 3 = SW solver. */
RRNN_API int rrnn_ctx_profile(rrnn_ctx_t* ctx, int enable);
RRNN_API int rrnn_ctx_profile_read(rrnn_ctx_t* ctx, int kind, double* total_ms, int* nlaunches);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
RRNN_API long long rrnn_ctx_launch_count(rrnn_ctx_t* ctx);
/* Which MLP kernel served the most recent NN gas-optics call on this context (the reference has ONE code path per model
 * set, predict_nn_{lw,sw}_blas_sp, rrtmgp/kernels/mo_gas_optics_kernels.F90:690-774, 869-953; here the tcgen05 kernel takes
 * every shipped model set and the fp32 FFMA kernel is the nn_tensor_cores = 0 variant and the fallback for anything else):
 * 0 none yet, RRNN_NN_KERNEL_FFMA, RRNN_NN_KERNEL_TCGEN05.  The counts are launches since the context was created. */
enum { RRNN_NN_KERNEL_NONE = 0, RRNN_NN_KERNEL_FFMA = 1, RRNN_NN_KERNEL_TCGEN05 = 2 };
RRNN_API int rrnn_ctx_last_nn_kernel(rrnn_ctx_t* ctx);
RRNN_API int rrnn_ctx_nn_kernel_counts(rrnn_ctx_t* ctx, long long* n_tcgen05, long long* n_ffma);
/* Device memory for hosts without a CUDA binding of their own (the Fortran veneer, fortran/mo_rrnn_veneer.F90): the derived
 * types of the reference own allocatable host arrays (tau/ssa/g, sources: rte/mo_optical_props.F90:98-192,
 * rte/mo_source_functions.F90:26-43); here they own device buffers.  Copies are ordered on the context's stream and complete
 * before the call returns. */
RRNN_API int rrnn_dev_malloc(rrnn_ctx_t* ctx, size_t bytes, void** out_d);
RRNN_API int rrnn_dev_free(rrnn_ctx_t* ctx, void* p_d);
RRNN_API int rrnn_memcpy_h2d(rrnn_ctx_t* ctx, void* dst_d, const void* src, size_t bytes);
RRNN_API int rrnn_memcpy_d2h(rrnn_ctx_t* ctx, void* dst, const void* src_d, size_t bytes);

/* ------------------------------------------------------------------------------------------------ */
/* NN models: rrtmgp_network_type%load_netcdf, neural/mod_network_rrtmgp.F90:58-122                   */
RRNN_API int rrnn_model_load_netcdf(rrnn_ctx_t* ctx, const char* filename, rrnn_model_t** out);
/* ASCII format of network_type%load (neural/mod_network.F90:163-209) + sidecar scaling file; see INTEGRATION.md */
RRNN_API int rrnn_model_load_ascii(rrnn_ctx_t* ctx, const char* model_txt, const char* scaling_txt, rrnn_model_t** out);
RRNN_API int rrnn_model_save_ascii(const rrnn_model_t* m, const char* model_txt, const char* scaling_txt);
/* Build from host arrays: wpack = layer weights back to back, each row-major (n_in,n_out) (== the
 * reference's column-major w_transposed(n_out,n_in)); input_names = nx*32 chars; ymean/ystd may be NULL. */
RRNN_API int rrnn_model_create(rrnn_ctx_t* ctx, int nlayers, const int* dims, const float* wpack, const float* bpack,
                               const int* activations, const float* xmin, const float* xmax, const float* ymean,
                               const float* ystd, const char* input_names, rrnn_model_t** out);
RRNN_API int rrnn_model_destroy(rrnn_model_t* m);
RRNN_API int rrnn_model_nlayers(const rrnn_model_t* m);
RRNN_API int rrnn_model_dims(const rrnn_model_t* m, int* dims_out /* nlayers+1 */);
RRNN_API int rrnn_model_input_name(const rrnn_model_t* m, int i, char* buf32);
RRNN_API int rrnn_model_activation(const rrnn_model_t* m, int layer);
/* host copies of the parameters (for cross-checking the reader): which = 0 weights(layer), 1 bias(layer),
 * 2 xmin, 3 xmax, 4 ymean, 5 ystd.  Returns the element count through n_out; data_out may be NULL. */
RRNN_API int rrnn_model_get(const rrnn_model_t* m, int which, int layer, float* data_out, int* n_out);

/* ------------------------------------------------------------------------------------------------ */
/* Spectral tables (ty_gas_optics_rrtmgp%load, rrtmgp/mo_gas_optics_rrtmgp.F90:1130-1326): band->g-point
 * limits (2,nbnd) 1-based inclusive, totplnk (nPlanckTemp,nbnd) == C [nbnd][ntemp] (may be NULL for SW),
 * solar_source (ngpt) (may be NULL for LW). */
RRNN_API int rrnn_kdist_create(rrnn_ctx_t* ctx, int nbnd, int ngpt, const int* band_lims_gpt, int ntemp,
                               const float* totplnk, float temp_ref_min, float totplnk_delta,
                               const float* solar_source, rrnn_kdist_t** out);
RRNN_API int rrnn_kdist_destroy(rrnn_kdist_t* kd);
/* ty_gas_optics_rrtmgp%set_tsi, rrtmgp/mo_gas_optics_rrtmgp.F90:1097-1120 */
RRNN_API int rrnn_kdist_set_tsi(rrnn_kdist_t* kd, float tsi);
/* The optional solar tables of ty_gas_optics_rrtmgp%load (load_ext, rrtmgp/mo_gas_optics_rrtmgp.F90:1317-1325):
 * solar_source_quiet / _facular / _sunspot, HOST arrays of ngpt entries. */
RRNN_API int rrnn_kdist_set_solar_tables(rrnn_kdist_t* kd, const float* solar_quiet, const float* solar_facular,
                                         const float* solar_sunspot);
/* ty_gas_optics_rrtmgp%set_solar_variability, rrtmgp/mo_gas_optics_rrtmgp.F90:1058-1095: solar_source = quiet +
 * (mg_index - 0.1495954) facular + (sb_index - 0.00066696) sunspot, then set_tsi(tsi) when have_tsi != 0. */
RRNN_API int rrnn_kdist_set_solar_variability(rrnn_kdist_t* kd, float mg_index, float sb_index, int have_tsi, float tsi);
/* ty_solar_var%solar_var_ind_interp, extensions/solar_variability/mo_solar_variability.F90:91-183: facular (mg) and sunspot (sb)
 * indices of the mean solar cycle interpolated to the cycle fraction solcycfrac in [0, 1] -- what set_solar_variability takes.
 * avgcyc_ind is the table ty_solar_var%load keeps (:45-69), Fortran (nsolarterms = 2, nsolarfrac) == C [nsolarfrac][2], a HOST
 * array; a host-only routine in the reference and here (no context, no device). */
RRNN_API int rrnn_solar_var_ind_interp(const float* avgcyc_ind, int nsolarfrac, float solcycfrac, float* mg_index_out, float* sb_index_out);
/* Host copy of the current solar source (ngpt). */
RRNN_API int rrnn_kdist_get_solar_source(const rrnn_kdist_t* kd, float* solar_source_out);
/* optimal_angle_fit (2,nbnd) of ty_gas_optics_rrtmgp%load (:1163, 1210), a HOST array == C [nbnd][2]. */
RRNN_API int rrnn_kdist_set_optimal_angle_fit(rrnn_kdist_t* kd, const float* optimal_angle_fit);
/* ty_gas_optics_rrtmgp%compute_optimal_angles, rrtmgp/mo_gas_optics_rrtmgp.F90:1712-1758: secant per column and g-point
 * from the column transmissivity, fit(1,bnd) exp(-sum_lay tau) + fit(2,bnd).  tau_d (ngpt,nlay,ncol); optimal_angles_d
 * (ngpt,ncol) -- the layout rte_lw's lw_Ds takes here (the reference declares (ncol,ngpt), a stale upstream order). */
RRNN_API int rrnn_compute_optimal_angles(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, const float* tau_d,
                                         float* optimal_angles_d);

/* ------------------------------------------------------------------------------------------------ */
/* Gas optics building blocks (device pointers)                                                       */
/* get_col_dry, rrtmgp/mo_gas_optics_rrtmgp.F90:1662-1707 (latitude absent) */
RRNN_API int rrnn_get_col_dry(rrnn_ctx_t* ctx, int ncol, int nlay, const float* vmr_h2o_d, const float* plev_d, float* col_dry_d);
/* level-temperature interpolation, rrtmgp/mo_gas_optics_rrtmgp.F90:326-335 */
RRNN_API int rrnn_interp_tlev(rrnn_ctx_t* ctx, int ncol, int nlay, const float* play_d, const float* plev_d, const float* tlay_d, float* tlev_d);
/* compute_nn_inputs, rrtmgp/mo_gas_optics_rrtmgp.F90:618-798 -> nn_inputs (ninputs,nlay,ncol) */
RRNN_API int rrnn_compute_nn_inputs(rrnn_ctx_t* ctx, const rrnn_model_t* m, int ncol, int nlay, const float* play_d,
                                    const float* tlay_d, const rrnn_gas_t* gases, int ngas, float* nn_inputs_d);
/* output_sgemm_tau / _pfrac / _lw, neural/mod_network_rrtmgp.F90:125-236, 238-317, 319-409 (x: (nx,nbatch)) */
RRNN_API int rrnn_output_sgemm_tau(rrnn_ctx_t* ctx, const rrnn_model_t* m, int nbatch, const float* x_d, const float* coldry_d,
                                   float* output_d, float* output2_d /* NULL or tau_abs in / tau_tot out */);
RRNN_API int rrnn_output_sgemm_pfrac(rrnn_ctx_t* ctx, const rrnn_model_t* m, int nbatch, const float* x_d, float* output_d);
RRNN_API int rrnn_output_sgemm_lw(rrnn_ctx_t* ctx, const rrnn_model_t* m, int nbatch, const float* x_d, float* output_d);
/* compute_Planck_source_nn, rrtmgp/kernels/mo_gas_optics_kernels.F90:615-683: pfrac_lay_source_d holds the
 * Planck fraction on input and lay_source on output; sfc_lay is 1-based. */
RRNN_API int rrnn_planck_source_nn(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int ncol, int nlay, const float* tlay_d,
                                   const float* tlev_d, const float* tsfc_d, int sfc_lay, float* sfc_source_d,
                                   float* sfc_source_Jac_d, float* pfrac_lay_source_d, float* lev_source_d);

/* ty_gas_optics_rrtmgp%gas_optics with neural_nets present.
 * LW = gas_optics_int NN branch, rrtmgp/mo_gas_optics_rrtmgp.F90:239-428 (:368-411): nmodels = 2 (tau net,
 *      Planck-fraction net) or 1 ("both" net, 2*ngpt outputs); tlev_d may be NULL (-> interpolation :326-335).
 *      One fused kernel: input scaling + col_dry + MLP chain + tau / Planck-source epilogues.
 * SW = gas_optics_ext NN branch, :433-602 (:529-573, :594-599): models[0] absorption, models[1] Rayleigh;
 *      ssa_d NULL -> 1scl request (absorption tau only); g_d NULL -> g (identically 0, :560-567) not materialised. */
RRNN_API int rrnn_gas_optics_lw(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                                int ncol, int nlay, const float* play_d, const float* plev_d, const float* tlay_d,
                                const float* tsfc_d, const rrnn_gas_t* gases, int ngas, const float* tlev_d,
                                float* tau_d, float* lay_source_d, float* lev_source_d, float* sfc_source_d,
                                float* sfc_source_Jac_d);
RRNN_API int rrnn_gas_optics_sw(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int ncol,
                                int nlay, const float* play_d, const float* plev_d, const float* tlay_d,
                                const rrnn_gas_t* gases, int ngas, float* tau_d, float* ssa_d, float* g_d,
                                float* toa_src_d);

/* The same longwave gas optics with the sources left FACTORED (no reference counterpart; this library's fused path
 * rrnn_lw_fluxes uses it): instead of lay_source(g,l) = pfrac(g,l) B_b(T_lay(l)) and lev_source (g,l) = pfrac(g,min(l,nlay))
 * B_b(T_lev(l)) (compute_Planck_source_nn, rrtmgp/kernels/mo_gas_optics_kernels.F90:654-672) it returns their factors:
 * pfrac_d (ngpt,nlay,ncol) and the band Planck functions planck_lay_d (16,nlay,ncol), planck_lev_d (16,nlay+1,ncol) (rows
 * of 16 floats, bands >= nbnd repeat the last band).  8 instead of 12 bytes per (g-point, layer) cross HBM.  Two networks,
 * tensor-core kernel only (an error otherwise); rrnn_lw_solver_noscat_compact consumes it with bit-identical fluxes. */
RRNN_API int rrnn_gas_optics_lw_compact(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models,
                                        int nmodels, int ncol, int nlay, const float* play_d, const float* plev_d,
                                        const float* tlay_d, const float* tsfc_d, const rrnn_gas_t* gases, int ngas,
                                        const float* tlev_d, float* tau_d, float* pfrac_d, float* planck_lay_d,
                                        float* planck_lev_d, float* sfc_source_d, float* sfc_source_Jac_d);

/* ------------------------------------------------------------------------------------------------ */
/* RTE solvers (device pointers)                                                                      */
/* lw_solver_noscat_GaussQuad / lw_solver_noscat, rte/kernels/mo_rte_solver_kernels.F90:332-415, 119-330
 * (no rescaling, no Jacobian): Ds/weights are HOST arrays of nmus entries; inc_flux_d may be NULL (= 0). */
RRNN_API int rrnn_lw_solver_noscat(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, int nmus, const float* Ds,
                                   const float* weights, const float* inc_flux_d, const float* tau_d,
                                   const float* lay_source_d, const float* lev_source_d, const float* sfc_emis_gpt_d,
                                   const float* sfc_source_d, float* flux_up_d, float* flux_dn_d);
/* lw_solver_noscat_GaussQuad (mo_rte_solver_kernels.F90:332-415) on the factored sources of rrnn_gas_optics_lw_compact:
 * the solver forms lay_source / lev_source itself (one fp32 product, as mo_gas_optics_kernels.F90:654-672).  kd supplies
 * the g-point -> band map.  Needs ngpt % 4 == 0, ngpt <= 512, nlay >= 8. */
RRNN_API int rrnn_lw_solver_noscat_compact(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int nmus,
                                           const float* Ds, const float* weights, const float* tau_d, const float* pfrac_d,
                                           const float* planck_lay_d, const float* planck_lev_d,
                                           const float* sfc_emis_gpt_d, const float* sfc_source_d, float* flux_up_d,
                                           float* flux_dn_d);
/* lw_solver_noscat_GaussQuad with every option rte_lw can pass it (rte/kernels/mo_rte_solver_kernels.F90:332-415, 119-330):
 * re-scaled scattering (do_rescaling: ssa_d, g_d, lw_transport_1rescl :1729-1795; both NULL = none), per-g-point secants
 * lw_Ds_gpt_d (ngpt,ncol; one angle, NULL = Ds[]), g-point fluxes gpt_flux_{up,dn}_d (ngpt,nlay+1,ncol; NULL = not wanted;
 * with one angle they hold radiances NOT multiplied by 2 pi w, as in the reference :287-291), and the surface-temperature
 * Jacobian flux_up_Jac_d (nlay+1,ncol) from sfc_source_Jac_d (compute_Jac, rte/mo_rte_rrtmgp_config.F90:29; with one
 * angle the sum of the un-scaled Jacobian radiances, :319).  A general kernel, not the tuned benchmark path. */
RRNN_API int rrnn_lw_solver_noscat_ext(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, int nmus, const float* Ds,
                                       const float* weights, const float* lw_Ds_gpt_d, const float* inc_flux_d,
                                       const float* tau_d, const float* ssa_d, const float* g_d, const float* lay_source_d,
                                       const float* lev_source_d, const float* sfc_emis_gpt_d, const float* sfc_source_d,
                                       const float* sfc_source_Jac_d, float* flux_up_d, float* flux_dn_d,
                                       float* flux_up_Jac_d, float* gpt_flux_up_d, float* gpt_flux_dn_d);
/* lw_solver_2stream (rte/kernels/mo_rte_solver_kernels.F90:426-486: lw_two_stream :1018-1069, lw_source_2str :1112-1162, adding
 * :1526-1637): two-stream longwave with scattering; lay_source is not used by the reference and is not an argument here;
 * gpt_flux_{up,dn}_d (ngpt,nlay+1,ncol) optional.  General kernel. */
RRNN_API int rrnn_lw_solver_2stream(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, const float* inc_flux_d,
                                    const float* tau_d, const float* ssa_d, const float* g_d, const float* lev_source_d,
                                    const float* sfc_emis_gpt_d, const float* sfc_source_d, float* flux_up_d,
                                    float* flux_dn_d, float* gpt_flux_up_d, float* gpt_flux_dn_d);
/* rte_lw(..., use_2stream = .true.) for ty_optical_props_2str (rte/mo_rte_lw.F90:346-361); sfc_emis_d is (nbnd,ncol). */
RRNN_API int rrnn_rte_lw_2stream(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1,
                                 const float* inc_flux_d, const float* tau_d, const float* ssa_d, const float* g_d,
                                 const float* lev_source_d, const float* sfc_source_d, const float* sfc_emis_d,
                                 float* flux_up_d, float* flux_dn_d, float* gpt_flux_up_d, float* gpt_flux_dn_d);
/* rte_lw with its optional arguments and for ty_optical_props_2str (re-scaled solution, rte/mo_rte_lw.F90:363-384):
 * ssa_d/g_d NULL = _1scl; lw_Ds_d (ngpt,ncol) only for _1scl and one angle (:239-249); sfc_emis_d is (nbnd,ncol). */
RRNN_API int rrnn_rte_lw_ext(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int n_gauss_angles,
                             const float* inc_flux_d, const float* tau_d, const float* ssa_d, const float* g_d,
                             const float* lay_source_d, const float* lev_source_d, const float* sfc_source_d,
                             const float* sfc_emis_d, const float* lw_Ds_d, const float* sfc_source_Jac_d, float* flux_up_d,
                             float* flux_dn_d, float* flux_up_Jac_d, float* gpt_flux_up_d, float* gpt_flux_dn_d);
/* rte_lw for ty_optical_props_1scl, rte/mo_rte_lw.F90:60-424: sfc_emis_d is (nbnd,ncol) and is expanded to
 * g-points (:429-447); n_gauss_angles in 1..4 with the secants/weights of :113-125. */
RRNN_API int rrnn_rte_lw(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int n_gauss_angles,
                         const float* inc_flux_d, const float* tau_d, const float* lay_source_d,
                         const float* lev_source_d, const float* sfc_source_d, const float* sfc_emis_d,
                         float* flux_up_d, float* flux_dn_d);
/* sw_solver_2stream, rte/kernels/mo_rte_solver_kernels.F90:541-692 (two-stream :1366-1480 + adding :1526-1637).
 * inc_flux_dif_d may be NULL (= 0, rte/mo_rte_sw.F90:191-205); g_d may be NULL (g = 0). */
RRNN_API int rrnn_sw_solver_2stream(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, const float* inc_flux_d,
                                    const float* inc_flux_dif_d, const float* tau_d, const float* ssa_d, const float* g_d,
                                    const float* mu0_d, const float* sfc_alb_dir_d, const float* sfc_alb_dif_d,
                                    float* flux_up_d, float* flux_dn_d, float* flux_dir_d);
/* sw_solver_2stream with its optional g-point fluxes (rte/kernels/mo_rte_solver_kernels.F90:541-692, save_gpt_flux):
 * gpt_flux_{up,dn,dir}_d (ngpt,nlay+1,ncol), all three or none; gpt_flux_dn is the TOTAL downward flux (:660-663).  The
 * reference's own three sweeps (sw_two_stream_source :1366-1480, adding :1526-1637), a general kernel, not the tuned one. */
RRNN_API int rrnn_sw_solver_2stream_ext(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, const float* inc_flux_d,
                                        const float* inc_flux_dif_d, const float* tau_d, const float* ssa_d, const float* g_d,
                                        const float* mu0_d, const float* sfc_alb_dir_d, const float* sfc_alb_dif_d,
                                        float* flux_up_d, float* flux_dn_d, float* flux_dir_d, float* gpt_flux_up_d,
                                        float* gpt_flux_dn_d, float* gpt_flux_dir_d);
/* rte_sw for ty_optical_props_2str, rte/mo_rte_sw.F90:48-266 (albedos per g-point, :50-61). */
RRNN_API int rrnn_rte_sw(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, const float* mu0_d,
                         const float* inc_flux_d, const float* sfc_alb_dir_gpt_d, const float* sfc_alb_dif_gpt_d,
                         const float* inc_flux_dif_d, const float* tau_d, const float* ssa_d, const float* g_d,
                         float* flux_up_d, float* flux_dn_d, float* flux_dn_dir_d);

/* rte_lw / rte_sw on gas optical properties + by-band cloud optical properties (nbnd,nlay,ncol) whose `clouds%increment(atmos)`
 * (rte/mo_optical_props.F90:714-893 -> inc_1scalar_by_1scalar_bybnd / inc_2stream_by_2stream_bybnd, rte/kernels/
 * mo_optical_props_kernels.F90:358-378, 453-485) has NOT been applied: the increment happens inside the solver, in registers,
 * instead of as a read-modify-write pass over tau / ssa / g.  SW: for gas properties with g == 0 (the NN gas optics); the cloud
 * properties are taken as given (delta-scale them first if the caller does, rrnn_delta_scale_2str on the by-band arrays).
 * Shapes the packed solvers do not take return an error that says so. */
RRNN_API int rrnn_rte_lw_clouds(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int n_gauss_angles,
                                const float* inc_flux_d, const float* tau_d, const float* lay_source_d, const float* lev_source_d,
                                const float* sfc_source_d, const float* sfc_emis_d, const float* cld_tau_bnd_d, float* flux_up_d,
                                float* flux_dn_d);
RRNN_API int rrnn_rte_sw_clouds(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, const float* mu0_d,
                                const float* inc_flux_d, const float* sfc_alb_dir_gpt_d, const float* sfc_alb_dif_gpt_d,
                                const float* inc_flux_dif_d, const float* tau_d, const float* ssa_d, const float* cld_tau_bnd_d,
                                const float* cld_ssa_bnd_d, const float* cld_g_bnd_d, float* flux_up_d, float* flux_dn_d,
                                float* flux_dn_dir_d);

/* ------------------------------------------------------------------------------------------------ */
/* Cloud optics (LUT), delta-scaling, increments, heating rates                                        */
/* ty_cloud_optics%load_lut, extensions/cloud_optics/mo_cloud_optics.F90:90-170: tables are
 * (nsize,nbnd) == C [nbnd][nsize] for the chosen ice roughness. */
RRNN_API int rrnn_cloud_lut_create(rrnn_ctx_t* ctx, int nbnd, int nsize_liq, int nsize_ice, float radliq_lwr, float radliq_upr,
                                   float radice_lwr, float radice_upr, const float* lut_extliq, const float* lut_ssaliq,
                                   const float* lut_asyliq, const float* lut_extice, const float* lut_ssaice,
                                   const float* lut_asyice, rrnn_cloud_lut_t** out);
/* load_pade (extensions/cloud_optics/mo_cloud_optics.F90:178-262): Pade coefficients as stored in the coefficient files,
 * (ncoeff, nsizereg, nbnd) with ncoeff_ext = 6 ([2/3]), ncoeff_ssa_g = 5 ([2/2]), nsizereg = 3, nbound = 4; the ice arrays for
 * the chosen roughness.  The handle is used with rrnn_cloud_optics like a LUT handle (compute_all_from_pade :650-714). */
RRNN_API int rrnn_cloud_pade_create(rrnn_ctx_t* ctx, int nbnd, int nsizereg, int ncoeff_ext, int ncoeff_ssa_g, int nbound,
                                    const float* pade_extliq, const float* pade_ssaliq, const float* pade_asyliq,
                                    const float* pade_extice, const float* pade_ssaice, const float* pade_asyice,
                                    const float* sizreg_extliq, const float* sizreg_ssaliq, const float* sizreg_asyliq,
                                    const float* sizreg_extice, const float* sizreg_ssaice, const float* sizreg_asyice,
                                    rrnn_cloud_lut_t** out);
RRNN_API int rrnn_cloud_lut_destroy(rrnn_cloud_lut_t* lut);
/* ty_cloud_optics%cloud_optics, mo_cloud_optics.F90:354-535 (compute_all_from_table :603-645): by-band
 * (nbnd,nlay,ncol) outputs; ssa_d = g_d = NULL -> 1scl absorption optical depth (:505-513). */
RRNN_API int rrnn_cloud_optics(rrnn_ctx_t* ctx, const rrnn_cloud_lut_t* lut, int ncol, int nlay, const float* clwp_d,
                               const float* ciwp_d, const float* reliq_d, const float* reice_d, float* tau_d, float* ssa_d,
                               float* g_d);
/* McICA sampling (extensions/cloud_optics/mo_cloud_sampling.F90): sampled_mask_max_ran :107-170 (overlap_param_d NULL) and
 * sampled_mask_exp_ran :176-286; randoms_d and the mask (1 byte per element) are (ngpt,nlay,ncol), cloud_frac_d (nlay,ncol),
 * overlap_param_d (nlay-1,ncol) -- this fork's layout throughout (the module itself still declares (ncol,nlay,ngpt)). */
RRNN_API int rrnn_sampled_mask(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, const float* randoms_d, const float* cloud_frac_d,
                               const float* overlap_param_d, unsigned char* cloud_mask_d);
/* draw_samples / apply_cloud_mask (:38-101, 292-308): by-band cloud properties (nbnd,nlay,ncol) -> sampled by g-point
 * (ngpt,nlay,ncol), zero where the mask is false; ssa/g NULL for ty_optical_props_1scl. */
RRNN_API int rrnn_draw_samples(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, const unsigned char* cloud_mask_d,
                               const float* tau_bnd_d, const float* ssa_bnd_d, const float* g_bnd_d, float* tau_gpt_d,
                               float* ssa_gpt_d, float* g_gpt_d);
/* delta_scale_2str_k, rte/kernels/mo_optical_props_kernels.F90:72-93 (n = number of elements) */
RRNN_API int rrnn_delta_scale_2str(rrnn_ctx_t* ctx, size_t n, float* tau_d, float* ssa_d, float* g_d);
/* inc_1scalar_by_1scalar_bybnd :358-378 and inc_2stream_by_2stream_bybnd :453-485; gpt_lims from kd.
 */
RRNN_API int rrnn_increment_1scl_bybnd(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, float* tau1_d, const float* tau2_d);
RRNN_API int rrnn_increment_2str_bybnd(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, float* tau1_d, float* ssa1_d,
                                       float* g1_d, const float* tau2_d, const float* ssa2_d, const float* g2_d);
/* compute_heating_rate, extensions/mo_heating_rates.F90:26-54 [K/s] (this fork's (nlay+1,ncol) layout) and
 * calc_heating_rate, examples/rrtmgp-nn-training/rrtmgp_lw_eval_nn_rfmip.F90:624-653 [K/day]. */
RRNN_API int rrnn_heating_rate(rrnn_ctx_t* ctx, int ncol, int nlay, const float* flux_up_d, const float* flux_dn_d,
                               const float* plev_d, float* heating_rate_d);
RRNN_API int rrnn_calc_heating_rate(rrnn_ctx_t* ctx, int ncol, int nlay, const float* flux_up_d, const float* flux_dn_d,
                                    const float* plev_d, float* hr_K_day_d);

/* ty_fluxes_byband%reduce, extensions/mo_fluxes_byband.F90:41-131.  sum_byband (mo_fluxes_byband_kernels.F90:33-51):
 * g-point fluxes (ngpt,nlev,ncol) -> by-band (nbnd,nlev,ncol), summed in g-point order; net_byband_full (:56-78):
 * by-band sum of (down - up).  Band limits from kd.  This fork's layout (g-point / band fastest). */
RRNN_API int rrnn_sum_byband(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlev, int ncol, const float* gpt_flux_d,
                             float* bnd_flux_d);
RRNN_API int rrnn_net_byband(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlev, int ncol, const float* gpt_flux_dn_d,
                             const float* gpt_flux_up_d, float* bnd_flux_net_d);
/* rte_lw / rte_sw with ty_fluxes_byband (extensions/mo_fluxes_byband.F90:41-131) WITHOUT g-point fluxes: broadband and by-band fluxes
 * (nbnd,nlay+1,ncol) both come out of the tuned solver, whose per-level sums pass through the band sums anyway.  Needs bands of 16
 * aligned g-points (every RRTMGP k-distribution) and a shape the packed solver takes; otherwise an error that says so, and the
 * general path (g-point fluxes + rrnn_sum_byband) is the one to use.  Summation order differs from sum_byband's (in g-point order),
 * so these agree with it to rounding, not bit for bit.  LW with ONE quadrature angle: the by-band values are sums of the reference's
 * un-scaled g-point radiances, as there (quirk Q3, rte/kernels/mo_rte_solver_kernels.F90:284-291); the broadband fluxes are fluxes. */
RRNN_API int rrnn_rte_lw_byband(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int n_gauss_angles,
                                const float* inc_flux_d, const float* tau_d, const float* lay_source_d, const float* lev_source_d,
                                const float* sfc_source_d, const float* sfc_emis_d, float* flux_up_d, float* flux_dn_d,
                                float* bnd_flux_up_d, float* bnd_flux_dn_d);
RRNN_API int rrnn_rte_sw_byband(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, const float* mu0_d,
                                const float* inc_flux_d, const float* sfc_alb_dir_gpt_d, const float* sfc_alb_dif_gpt_d,
                                const float* inc_flux_dif_d, const float* tau_d, const float* ssa_d, const float* g_d, float* flux_up_d,
                                float* flux_dn_d, float* flux_dn_dir_d, float* bnd_flux_up_d, float* bnd_flux_dn_d,
                                float* bnd_flux_dn_dir_d);
/* net = down - up over n elements: net_byband_precalc (mo_fluxes_byband_kernels.F90:80-86) and the broadband flux_net of
 * ty_fluxes_broadband%reduce (rte/mo_fluxes.F90, net_broadband_precalc). */
RRNN_API int rrnn_net_flux(rrnn_ctx_t* ctx, size_t n, const float* flux_dn_d, const float* flux_up_d, float* flux_net_d);

/* ------------------------------------------------------------------------------------------------ */
/* Whole-path drivers with HOST buffers: what one iteration of the reference drivers' block loop does
 * (examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:368-446, rrtmgp_rfmip_sw.F90:356-465): gas_optics -> rte_lw /
 * rte_sw, here for all columns at once, in column chunks that fit the device workspace, H2D/D2H overlapped
 * with compute.  sfc_emis/sfc_alb are per column (spectrally constant, as in the RFMIP drivers); mu0 <= 0
 * marks night columns (fluxes zeroed, rrtmgp_rfmip_sw.F90:458-463).  tsi_scale may be NULL; else the
 * per-column TSI renormalisation of :409-416 is applied.  Gas conc pointers are HOST pointers.           */
RRNN_API int rrnn_lw_fluxes_host(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                                 int ncol, int nlay, int top_at_1, int n_gauss_angles, const float* play, const float* plev,
                                 const float* tlay, const float* tlev, const float* tsfc, const float* sfc_emis,
                                 const rrnn_gas_t* gases, int ngas, float* flux_up, float* flux_dn);
RRNN_API int rrnn_sw_fluxes_host(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int ncol,
                                 int nlay, int top_at_1, const float* play, const float* plev, const float* tlay,
                                 const float* mu0, const float* sfc_alb, const float* tsi, const rrnn_gas_t* gases,
                                 int ngas, float* flux_up, float* flux_dn, float* flux_dn_dir);
/* Same path with DEVICE buffers (inputs already resident; used for the kernel-only throughput number). */
RRNN_API int rrnn_lw_fluxes(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                            int ncol, int nlay, int top_at_1, int n_gauss_angles, const float* play_d, const float* plev_d,
                            const float* tlay_d, const float* tlev_d, const float* tsfc_d, const float* sfc_emis_d,
                            const rrnn_gas_t* gases, int ngas, float* flux_up_d, float* flux_dn_d);
RRNN_API int rrnn_sw_fluxes(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int ncol, int nlay,
                            int top_at_1, const float* play_d, const float* plev_d, const float* tlay_d, const float* mu0_d,
                            const float* sfc_alb_d, const float* tsi_d, const rrnn_gas_t* gases, int ngas,
                            float* flux_up_d, float* flux_dn_d, float* flux_dn_dir_d);
/* All-sky whole-path drivers: one iteration of examples/all-sky/rrtmgp_allsky.F90:366-446 -- cloud_optics (LUT or Pade handle, by
 * band), gas_optics(neural_nets=), [delta_scale,] increment, rte_lw / rte_sw -- for all columns of the call.  clwp / ciwp /
 * reliq / reice are (nlay,ncol); the other arguments are those of the clear-sky drivers above.  The cloud increment is NOT a
 * pass over the (ngpt,nlay,ncol) arrays: the by-band cloud properties are added to the gas properties inside the solvers. */
RRNN_API int rrnn_lw_fluxes_allsky(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                                   const rrnn_cloud_lut_t* cloud_optics, int ncol, int nlay, int top_at_1, int n_gauss_angles,
                                   const float* play_d, const float* plev_d, const float* tlay_d, const float* tlev_d, const float* tsfc_d,
                                   const float* sfc_emis_d, const rrnn_gas_t* gases, int ngas, const float* clwp_d, const float* ciwp_d,
                                   const float* reliq_d, const float* reice_d, float* flux_up_d, float* flux_dn_d);
RRNN_API int rrnn_sw_fluxes_allsky(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models,
                                   const rrnn_cloud_lut_t* cloud_optics, int ncol, int nlay, int top_at_1, const float* play_d,
                                   const float* plev_d, const float* tlay_d, const float* mu0_d, const float* sfc_alb_d, const float* tsi_d,
                                   const rrnn_gas_t* gases, int ngas, const float* clwp_d, const float* ciwp_d, const float* reliq_d,
                                   const float* reice_d, float* flux_up_d, float* flux_dn_d, float* flux_dn_dir_d);
RRNN_API int rrnn_lw_fluxes_allsky_host(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                                        const rrnn_cloud_lut_t* cloud_optics, int ncol, int nlay, int top_at_1, int n_gauss_angles,
                                        const float* play, const float* plev, const float* tlay, const float* tlev, const float* tsfc,
                                        const float* sfc_emis, const rrnn_gas_t* gases, int ngas, const float* clwp, const float* ciwp,
                                        const float* reliq, const float* reice, float* flux_up, float* flux_dn);
RRNN_API int rrnn_sw_fluxes_allsky_host(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models,
                                        const rrnn_cloud_lut_t* cloud_optics, int ncol, int nlay, int top_at_1, const float* play,
                                        const float* plev, const float* tlay, const float* mu0, const float* sfc_alb, const float* tsi,
                                        const rrnn_gas_t* gases, int ngas, const float* clwp, const float* ciwp, const float* reliq,
                                        const float* reice, float* flux_up, float* flux_dn, float* flux_dn_dir);
/* Column chunk used by the drivers above (0 = automatic from free device memory). */
RRNN_API int rrnn_ctx_set_chunk_columns(rrnn_ctx_t* ctx, int ncol_chunk);

/* ------------------------------------------------------------------------------------------------ */
/* netCDF-4 files on either side of the path (no libnetcdf / libhdf5 needed; csrc/nc4_io.cpp, csrc/nc4.hpp): the calls the
 * reference's drivers make through examples/mo_simple_netcdf.F90 -- read_field :34-96, var_exists :308-318, create_dim :320-343,
 * create_var :345-380, write_field :167-237 -- and nf90_get_att for the RFMIP `units` scaling factors (examples/rfmip-clear-sky/
 * mo_rfmip_io.F90:560-600).  A handle is either open for reading (rrnn_nc_open) or a file being built (rrnn_nc_create; written by
 * rrnn_nc_close).  Arrays are in file order (C order of the netCDF dimensions). */
typedef struct rrnn_ncfile rrnn_ncfile_t;
RRNN_API int rrnn_nc_open(const char* path, rrnn_ncfile_t** out);
RRNN_API int rrnn_nc_create(const char* path, rrnn_ncfile_t** out);
RRNN_API int rrnn_nc_close(rrnn_ncfile_t* f);
RRNN_API int rrnn_nc_var_exists(const rrnn_ncfile_t* f, const char* name);   /* 1 / 0 */
RRNN_API int rrnn_nc_inq_var(const rrnn_ncfile_t* f, const char* name, int* ndims, long long* shape /* [8] */);
/* any numeric variable, converted to float (real(wp) read_field) */
RRNN_API int rrnn_nc_get_var_float(const rrnn_ncfile_t* f, const char* name, float* data_out, size_t n);
RRNN_API int rrnn_nc_get_att_text(const rrnn_ncfile_t* f, const char* var, const char* att, char* buf, int nbuf);
RRNN_API int rrnn_nc_def_dim(rrnn_ncfile_t* f, const char* name, long long len, int* dimid);
/* create_var + write_field: a float variable over already-defined dimensions; units may be NULL */
RRNN_API int rrnn_nc_put_var_float(rrnn_ncfile_t* f, const char* name, int ndims, const int* dimids, const float* data,
                                   const char* units);

/* ------------------------------------------------------------------------------------------------ */
/* One process, N devices.  The reference's drivers run their column blocks under an OpenMP parallel-do with firstprivate
 * copies of the k-distribution and the networks (examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:364-368, rrtmgp_rfmip_sw.F90:
 * 352-356); here the "threads" are GPUs: a rrnn_multi_t holds one context per device with the spectral tables and the networks
 * replicated on each, the columns of a call are cut into contiguous shards (sizes differing by at most one column), every
 * shard runs rrnn_{lw,sw}_fluxes_host on its own host thread and writes its fluxes straight into its slice of the caller's
 * HOST arrays.  No exchange step, hence no collective.  devices = NULL means devices 0 .. ndev-1; a device may be listed twice. */
typedef struct rrnn_multi rrnn_multi_t;
RRNN_API int rrnn_multi_create(int ndev, const int* devices, rrnn_multi_t** out);
RRNN_API int rrnn_multi_destroy(rrnn_multi_t* m);
RRNN_API int rrnn_multi_ndev(const rrnn_multi_t* m);
RRNN_API rrnn_ctx_t* rrnn_multi_ctx(rrnn_multi_t* m, int i);               /* the i-th device's context (flags, profiling) */
RRNN_API int rrnn_multi_set_flag(rrnn_multi_t* m, const char* name, int value);   /* rrnn_ctx_set_flag on every device */
/* rrtmgp_network_type%load_netcdf (neural/mod_network_rrtmgp.F90:58-122) onto every device; returns an id for the calls below */
RRNN_API int rrnn_multi_model_load_netcdf(rrnn_multi_t* m, const char* filename, int* model_id);
/* rrnn_kdist_create on every device; returns an id */
RRNN_API int rrnn_multi_kdist_create(rrnn_multi_t* m, int nbnd, int ngpt, const int* band_lims_gpt, int ntemp, const float* totplnk,
                                     float temp_ref_min, float totplnk_delta, const float* solar_source, int* kdist_id);
RRNN_API int rrnn_multi_kdist_set_tsi(rrnn_multi_t* m, int kdist_id, float tsi);
/* rrnn_lw_fluxes_host / rrnn_sw_fluxes_host over all devices (same arguments, ids instead of handles) */
RRNN_API int rrnn_multi_lw_fluxes_host(rrnn_multi_t* m, int kdist_id, const int* model_ids, int nmodels, int ncol, int nlay, int top_at_1,
                                       int n_gauss_angles, const float* play, const float* plev, const float* tlay, const float* tlev,
                                       const float* tsfc, const float* sfc_emis, const rrnn_gas_t* gases, int ngas, float* flux_up,
                                       float* flux_dn);
RRNN_API int rrnn_multi_sw_fluxes_host(rrnn_multi_t* m, int kdist_id, const int* model_ids, int ncol, int nlay, int top_at_1,
                                       const float* play, const float* plev, const float* tlay, const float* mu0, const float* sfc_alb,
                                       const float* tsi, const rrnn_gas_t* gases, int ngas, float* flux_up, float* flux_dn,
                                       float* flux_dn_dir);

#ifdef __cplusplus
}
#endif
#endif /* RRNN_H */
