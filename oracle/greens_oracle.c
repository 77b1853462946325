/*
 * oracle/greens_oracle.c -- CPU restatement of the reference's Green's-function
 * kernels.  TEST INFRASTRUCTURE ONLY: nothing under verde_b200/ may link, load
 * or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / CPU yardstick.
 *
 * Each function restates one numba kernel of fatiando/verde (loops, operation
 * order and branch structure kept; threads = OpenMP over the same index the
 * reference gives to numba.prange):
 *
 *   oracle_greens            <- verde/spline.py:587-605  greens_func_jit
 *   oracle_jacobian          <- verde/spline.py:642-650  jacobian_numba
 *   oracle_predict           <- verde/spline.py:629-639  predict_numba
 *   oracle_greens_2d         <- verde/vector.py:393-405  greens_func_2d
 *   oracle_jacobian_2d       <- verde/vector.py:461-475  jacobian_2d_numba
 *   oracle_predict_2d        <- verde/vector.py:443-458  predict_2d_numba
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp, NO -ffast-math: the scalar
 * reference kernels are compiled without fastmath; the vector ones use numba
 * fastmath=True, which only licenses reassociation, so the strict evaluation
 * here is one of the orders fastmath may pick).
 */
#include <math.h>
#include <stddef.h>

/* verde/spline.py:587-605 */
double oracle_greens(double east, double north, double mindist)
{
    double distance = sqrt(east * east + north * north);
    distance += mindist;
    if (distance < 1.0) {
        /* verde/spline.py:602  distance * (log(distance**distance) - distance) */
        return distance * (log(pow(distance, distance)) - distance);
    }
    /* verde/spline.py:604  distance**2 * (log(distance) - 1) */
    return distance * distance * (log(distance) - 1.0);
}

/* verde/spline.py:642-650: jac is row-major (n_data, n_forces), caller-allocated */
void oracle_jacobian(const double *east, const double *north, size_t n,
                     const double *force_east, const double *force_north, size_t m,
                     double mindist, double *jac)
{
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) {
        for (size_t j = 0; j < m; ++j) {
            jac[(size_t)i * m + j] =
                oracle_greens(east[i] - force_east[j], north[i] - force_north[j], mindist);
        }
    }
}

/* verde/spline.py:629-639: result[i] accumulated over j ascending */
void oracle_predict(const double *east, const double *north, size_t n,
                    const double *force_east, const double *force_north,
                    const double *forces, size_t m, double mindist, double *result)
{
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) {
        double acc = 0.0;
        for (size_t j = 0; j < m; ++j) {
            double green =
                oracle_greens(east[i] - force_east[j], north[i] - force_north[j], mindist);
            acc += green * forces[j];
        }
        result[i] = acc;
    }
}

/* verde/vector.py:393-405 */
void oracle_greens_2d(double east, double north, double mindist, double poisson,
                      double *green_ee, double *green_nn, double *green_ne)
{
    double distance = sqrt(east * east + north * north);
    distance += mindist;
    double ln_r = (3.0 - poisson) * log(distance);
    double over_r2 = (1.0 + poisson) / (distance * distance);
    *green_ee = ln_r + over_r2 * (north * north);
    *green_nn = ln_r + over_r2 * (east * east);
    *green_ne = -over_r2 * east * north;
}

/* verde/vector.py:461-475: jac is row-major (2n, 2m) */
void oracle_jacobian_2d(const double *east, const double *north, size_t n,
                        const double *force_east, const double *force_north, size_t m,
                        double mindist, double poisson, double *jac)
{
    const size_t ld = 2 * m;
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) {
        for (size_t j = 0; j < m; ++j) {
            double ee, nn, ne;
            oracle_greens_2d(east[i] - force_east[j], north[i] - force_north[j], mindist,
                             poisson, &ee, &nn, &ne);
            jac[(size_t)i * ld + j] = ee;
            jac[((size_t)i + n) * ld + j + m] = nn;
            jac[(size_t)i * ld + j + m] = ne;
            jac[((size_t)i + n) * ld + j] = ne;
        }
    }
}

/* verde/vector.py:443-458: forces = [east forces (m), north forces (m)] */
void oracle_predict_2d(const double *east, const double *north, size_t n,
                       const double *force_east, const double *force_north,
                       const double *forces, size_t m, double mindist, double poisson,
                       double *vec_east, double *vec_north)
{
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) {
        double ve = 0.0, vn = 0.0;
        for (size_t j = 0; j < m; ++j) {
            double ee, nn, ne;
            oracle_greens_2d(east[i] - force_east[j], north[i] - force_north[j], mindist,
                             poisson, &ee, &nn, &ne);
            ve += ee * forces[j] + ne * forces[j + m];
            vn += ne * forces[j] + nn * forces[j + m];
        }
        vec_east[i] = ve;
        vec_north[i] = vn;
    }
}
