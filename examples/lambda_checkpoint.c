/* lambda_checkpoint.c — a C host of libvrt.so: NLTE Λ-iteration on a Voronoi grid with the reference's per-iteration
 * checkpoint (Λ_voronoi, reference src/lambda_iteration.jl:253-285: after every iteration write_to_file(S_λ) and
 * write_to_file(populations), io.jl:57-83, and the convergence value, io.jl:129-135) done by the library from the device.
 *
 *   cc -I include examples/lambda_checkpoint.c -L voronoirt_b200 -lvrt -Wl,-rpath,$PWD/voronoirt_b200 -o lambda_checkpoint
 *
 * The caller supplies what compare_line.jl has at this point: positions (3 x n, rows z, x, y), the voro++ neighbour file, the
 * per-site arrays and the line constants.  Error handling: every entry returns 0 or a negative VRT_E_*; vrt_last_error() has
 * the message. */
#include <stdio.h>
#include <stdlib.h>
#include "vrt.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != VRT_OK) {                                                          \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, vrt_last_error());    \
            return rc_;                                                               \
        }                                                                             \
    } while (0)

struct checkpoint {
    vrt_outfile* file;
    vrt_solver* solver;
};

/* vrt_iter_cb: called once per Λ-iteration; a non-zero return stops the loop */
static int on_iteration(const vrt_iter_info* info, void* user) {
    struct checkpoint* c = (struct checkpoint*)user;
    if (vrt_output_write_state(c->file, c->solver) != VRT_OK) return 1;                       /* S_λ and the populations */
    if (vrt_output_write_convergence(c->file, info->iteration, info->diff) != VRT_OK) return 1;
    printf("iteration %d: diff %.3e, sweep %.1f ms, opacity %.1f ms\n", info->iteration, info->diff, info->t_sweep_ms, info->t_opacity_ms);
    return 0;
}

int run_lambda(const char* neighbour_file, int64_t n, const double* positions, const double bounds[6], const vrt_line* line,
               const double* lambda, const vrt_site_data* sites, const vrt_quadrature* quad, const char* output_path) {
    /* read_cell: neighbour text file -> NeighbourMatrix -> grid (voronoi_utils.jl:36-85) */
    int64_t ld = 0;
    CHECK(vrt_read_neighbours(neighbour_file, n, NULL, 0, &ld));
    int64_t* nbr = (int64_t*)calloc((size_t)n * (size_t)ld, sizeof(int64_t));
    if (!nbr) return VRT_E_NOMEM;
    CHECK(vrt_read_neighbours(neighbour_file, n, nbr, ld, NULL));
    vrt_grid* grid = NULL;
    CHECK(vrt_grid_create(n, positions, nbr, ld, bounds, &grid));
    free(nbr);

    vrt_config cfg = {0};
    cfg.n_sweeps = 3;
    cfg.p = 7.0;
    cfg.prune = 1;
    vrt_solver* solver = NULL;
    CHECK(vrt_solver_create_line(grid, line, lambda, sites, quad, &cfg, &solver));

    /* create_output_file(output_path, nλ, n_sites, maxiter) and the one-time write_to_file calls (io.jl:107-157, 196-225) */
    const int32_t maxiter = 150;
    struct checkpoint c = {NULL, solver};
    CHECK(vrt_output_create(output_path, line->nlam, n, maxiter, &c.file));
    CHECK(vrt_output_write(c.file, "positions", positions, (int64_t)sizeof(double) * 3 * n));
    CHECK(vrt_output_write(c.file, "temperature", sites->temperature, (int64_t)sizeof(double) * n));
    CHECK(vrt_output_write(c.file, "boundaries", bounds, (int64_t)sizeof(double) * 6));
    CHECK(vrt_output_write(c.file, "wavelength", lambda, (int64_t)sizeof(double) * line->nlam));
    CHECK(vrt_output_write(c.file, "line_center", &line->lambda0, (int64_t)sizeof(double)));

    vrt_result res;
    CHECK(vrt_lambda_iterate(solver, 1e-3, maxiter, on_iteration, &c, &res));
    printf("%s after %d iterations (diff %.3e, %.1f s)\n", res.converged ? "converged" : "stopped", res.iterations, res.diff, res.seconds);

    double t = res.seconds;
    CHECK(vrt_output_write(c.file, "time", &t, (int64_t)sizeof(double)));
    CHECK(vrt_output_close(c.file));
    vrt_solver_destroy(solver);
    vrt_grid_destroy(grid);
    return VRT_OK;
}
