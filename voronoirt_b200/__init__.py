"""voronoirt_b200 — B200-native (sm_100a) short-characteristics formal solver and Λ-iteration engine for
VoronoiRT's irregular-grid path, behind the reference's own function names.

    from voronoirt_b200 import read_cell, VoronoiSites, Delaunay_upII, Delaunay_downII, J_λ_voronoi,
                               calculate_R, get_revised_populations, Λ_voronoi

The compute lives in the in-tree C-ABI library voronoirt_b200/libvrt.so (include/vrt.h); this package is the
host-side mirror of the Julia interface.  There is no CPU fallback.
"""
from .api import (Atmosphere, J_lambda_regular, J_λ_regular, Lambda_regular, Λ_regular, Delaunay_downII, Delaunay_upII, J_lambda_voronoi, J_λ_voronoi, Lambda_voronoi, Solver,  # noqa: F401
                  VoronoiSites, calculate_R, direction, get_revised_populations, quadrature_path, read_cell,
                  read_neighbours, read_quadrature, regular_release_workspace, short_characteristics_down, voronoi_neighbours, trilinear, initialise, nearest_site, nearest_sites, Voronoi_to_Raster, Voronoi_to_Raster_inv_dist, initialiseII, rejection_sampling, short_characteristics_up, voro,
                  write_arrays, Λ_voronoi)
from .atom import B_λ, HydrogenicLine, LTE_populations, test_atom  # noqa: F401

__all__ = ["Atmosphere", "J_λ_regular", "J_lambda_regular", "Λ_regular", "Lambda_regular", "short_characteristics_up", "short_characteristics_down", "regular_release_workspace", "voronoi_neighbours", "trilinear", "initialise", "nearest_site", "nearest_sites", "Voronoi_to_Raster", "Voronoi_to_Raster_inv_dist", "initialiseII", "rejection_sampling", "Delaunay_downII", "Delaunay_upII", "J_lambda_voronoi", "J_λ_voronoi", "Lambda_voronoi", "Solver", "VoronoiSites",
           "calculate_R", "direction", "get_revised_populations", "quadrature_path", "read_cell", "read_neighbours",
           "read_quadrature", "voro", "write_arrays", "Λ_voronoi", "B_λ", "HydrogenicLine", "LTE_populations", "test_atom"]
