// qs_model.hpp — host-side construction of a slider model: .ply outline ingest, knot vector,
// derivative coefficient tables and the per-span polynomial (pp-form) tables the kernels evaluate.
//
// Follows /root/reference/acados_nmpc/PusherSliderModel.m:45-60 (constants), :84-111
// (sortCadPoints), :113-132 (getSpline) and bspline_shape.m:25-38, 85-104, 118-135 (coefficient
// tables).  The pp-form tables are obtained by running the reference's Cox-de Boor recursion
// (eval_bspline.m:1-33) in polynomial arithmetic on every knot span, in long double, so the
// kernels evaluate exactly the piecewise polynomials the reference's CasADi expressions define.
#pragma once
#include <string>
#include <vector>

#include "qs_device.cuh"

namespace qs {

struct HostModel {
    int p = 3, n = 0;
    std::vector<double> S, P, c1, c2;     // knots (n+p+1), control points n x 2, cj_1 n x 2, cj_2 n x 2
    double b = 0.0, mu_sp = 0.0, c_ellipse = 0.0;
    bool single_coeffs = true;
    std::vector<double> blob;             // MODEL_DOUBLES, layout in qs_device.cuh
};

// returns empty string on success, else an error message
std::string model_from_tables(const double* S, int nknots, const double* P, int n, int p, double mu_sp,
                              double c_ellipse, bool single_coeffs, HostModel& out);
std::string model_from_ply(const char* path, bool flip, int p, double mu_sg, double mu_sp, double mass,
                           double tau_max, HostModel& out);

}  // namespace qs
