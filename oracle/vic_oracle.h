/* vic_oracle.h -- ORACLE = TEST INFRASTRUCTURE.  Not part of the product.
 *
 * Plain-C CPU restatement of the reference's per-cell water-balance step
 * (runoff/model full_energy.c + surface_fluxes.c + ... + put_data.c; every function in
 * vic_oracle.c cites the reference file:line it follows).  It is pinned against the compiled
 * reference itself (oracle/_ref, tests/test_oracle_vs_ref.py) and against the committed golden
 * fixtures under tests/golden/.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product path (vicres_b200/csrc) never does.
 *
 * It takes the same plain-data structs as the product's C ABI (include/vicres.h) so that a
 * test can hand identical inputs to both sides.
 */
#ifndef VIC_ORACLE_H
#define VIC_ORACLE_H
#include "../include/vicres.h"

#ifdef __cplusplus
extern "C" {
#endif

#define VO_NUNITF 48   /* per tile x band debug fields, same order as oracle/refdump.py UNIT_FIELDS */

typedef struct vo_model vo_model;

vo_model *vo_create(const vr_options *opt, const vr_veglib *lib, const vr_cells *cells);
void      vo_destroy(vo_model *m);
int       vo_init_state(vo_model *m, const double *tair0);
/* out: [n_out][nrec][C]; units_dbg (may be NULL): [nrec][C][maxtile][nbands][VO_NUNITF] */
int       vo_step_block(vo_model *m, const vr_forcing *f, double *out, double *units_dbg, int maxtile);
/* advance cells [c0,c1) only (used by the multi-threaded cpu_baseline leg); thread-safe for
 * disjoint ranges */
int       vo_step_block_range(vo_model *m, const vr_forcing *f, double *out, int64_t c0, int64_t c1);
/* fallback counters (put_data.c:658-664): [0] = Tsnowsurf, [1] = Tfoliage */
void      vo_fallback_counts(const vo_model *m, int64_t counts[2]);

#ifdef __cplusplus
}
#endif
#endif
