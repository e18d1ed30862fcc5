"""A small regular, periodic atmosphere in the synthetic box (shared by the regular-grid line tests)."""
import numpy as np


def regular_line_box(oracle, nz=12, nx=10, ny=9, nbb=10, nbf=4):
    """-> dict with axes (ghost columns included), (nz, nx, ny) fields, the line inputs and the oracle structs.
    Cells are flattened column-major (z fastest), the memory order of the reference's arrays."""
    from voronoirt_b200 import synth
    B = synth.BOX
    z = np.linspace(B["z_min"], B["z_max"], nz) + np.linspace(0, 1, nz) ** 2 * 0.0
    z = B["z_min"] + (B["z_max"] - B["z_min"]) * np.linspace(0, 1, nz) ** 1.5          # stretched: the branch varies with height
    x = (np.arange(nx) - 1) * (B["x_max"] / (nx - 2))
    y = (np.arange(ny) - 1) * (B["y_max"] / (ny - 2))
    Z, X, Y = np.meshgrid(z, x, y, indexing="ij")
    a = synth.atmosphere(Z.ravel(order="F"), np.mod(X.ravel(order="F"), B["x_max"]), np.mod(Y.ravel(order="F"), B["y_max"]))
    f = {}
    for key in ("temperature", "electron_density", "hydrogen_density", "velocity_z", "velocity_x", "velocity_y"):
        v = np.asfortranarray(np.asarray(a[key], dtype=np.float64).reshape((nz, nx, ny), order="F"))
        v[:, 0, :] = v[:, -2, :]; v[:, -1, :] = v[:, 1, :]
        v[:, :, 0] = v[:, :, -2]; v[:, :, -1] = v[:, :, 1]
        f[key] = v
    flat = {k: v.ravel(order="F") for k, v in f.items()}
    line, lte, α_cont, ελ, Cr = synth.line_inputs(flat["temperature"], flat["electron_density"], flat["hydrogen_density"], nbb, nbf)
    sd = oracle.make_site_data(temperature=flat["temperature"], electron_density=flat["electron_density"],
                               hydrogen_density=flat["hydrogen_density"], velocity_z=flat["velocity_z"], velocity_x=flat["velocity_x"],
                               velocity_y=flat["velocity_y"], doppler_width=line.ΔD, alpha_cont=α_cont, destruction=ελ,
                               C=np.ascontiguousarray(Cr.T), lte_pops=np.ascontiguousarray(lte.T))
    return dict(z=z, x=x, y=y, shape=(nz, nx, ny), fields=f, flat=flat, line=line, lte=lte, α_cont=α_cont, ελ=ελ, C=Cr, sd=sd,
                n=nz * nx * ny)
