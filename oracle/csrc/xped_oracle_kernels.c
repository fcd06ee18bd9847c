/* TEST / BASELINE INFRASTRUCTURE ONLY -- compiled hot loop of the CPU oracle's permute (never linked into the product library).
 *
 * Restates the data movement of Tensor::permute_impl (src/Core/Tensor.cpp:83-214): for every source tree pair, copy the sub-block out of its
 * sector matrix, shuffle it (PlainInterface::shuffle -> Eigen::Tensor::shuffle, src/Interfaces/TensorInterface_Eigen_impl.cpp:273-278) and add
 * coeff * shuffled block into the target sub-block (add_to_block_from_tensor, src/Interfaces/PlainInterface_Eigen_impl.cpp:76-78).  The (source, target,
 * coefficient) list is produced once per (tensor structure, shift, permutation) by the Python oracle -- the reference's optional
 * XPED_CACHE_PERMUTE_OUTPUT LRU cache of treepair::permute (src/Core/treepair.cpp:170-176,211-215) plays the same role -- so the timed CPU path is
 * "cached recoupling coefficients + compiled strided transposes under OpenMP + BLAS/LAPACK", i.e. what an optimised Eigen+OpenMP build does.
 *
 * Record layout (int64, REC = 4 + 2 * MAXL per item): src_off, dst_off, nleg, coeff (double bits), dims[MAXL] (destination leg order, leg 0 fastest in
 * the destination), src_stride[MAXL], and the destination strides follow from the destination sub-block: dst_stride[MAXL] in a second table.
 * Items are grouped by destination sub-block: group g owns items [gstart[g], gstart[g+1]); groups write disjoint memory and run in parallel.
 */
#include <stdint.h>
#include <string.h>

#define XO_MAXL 8
#define XO_REC (4 + 3 * XO_MAXL)

static void one_item(const double* src, double* dst, const int64_t* r)
{
    const int nleg = (int)r[2];
    double coeff;
    memcpy(&coeff, &r[3], sizeof(double));
    const int64_t* dims = r + 4;
    const int64_t* ss = r + 4 + XO_MAXL;
    const int64_t* ds = r + 4 + 2 * XO_MAXL;
    const double* s0 = src + r[0];
    double* d0 = dst + r[1];
    if(nleg == 0) {
        d0[0] += coeff * s0[0];
        return;
    }
    /* legs 0 and 1 as plain nested loops, the remaining legs by an odometer */
    int64_t idx[XO_MAXL] = {0};
    const int64_t n0 = dims[0], ss0 = ss[0], ds0 = ds[0];
    const int64_t n1 = nleg > 1 ? dims[1] : 1, ss1 = nleg > 1 ? ss[1] : 0, ds1 = nleg > 1 ? ds[1] : 0;
    for(;;) {
        int64_t so = 0, dof = 0;
        for(int k = 2; k < nleg; ++k) {
            so += idx[k] * ss[k];
            dof += idx[k] * ds[k];
        }
        for(int64_t j = 0; j < n1; ++j) {
            const double* sp = s0 + so + j * ss1;
            double* dp = d0 + dof + j * ds1;
            if(ds0 == 1 && ss0 == 1)
                for(int64_t i = 0; i < n0; ++i) dp[i] += coeff * sp[i];
            else if(ds0 == 1)
                for(int64_t i = 0; i < n0; ++i) dp[i] += coeff * sp[i * ss0];
            else
                for(int64_t i = 0; i < n0; ++i) dp[i * ds0] += coeff * sp[i * ss0];
        }
        int k = 2;
        for(; k < nleg; ++k) {
            if(++idx[k] < dims[k]) break;
            idx[k] = 0;
        }
        if(k >= nleg) break;
    }
}

/* The reference's permute loop is serial (only the BLAS / LAPACK calls of an Eigen build are threaded); more threads here also fight with the BLAS thread
 * pool of the surrounding numpy process.  Default 1, xo_set_threads() for experiments. */
static int xo_threads = 1;
void xo_set_threads(int n) { xo_threads = n < 1 ? 1 : n; }

/* dst must be zero-initialised by the caller (Tensor::setZero, src/Core/Tensor.cpp:46-55) */
void xo_permute_apply(const double* src, double* dst, const int64_t* records, const int64_t* gstart, int64_t ngroups)
{
#pragma omp parallel for schedule(dynamic, 16) num_threads(xo_threads) if(xo_threads > 1)
    for(int64_t g = 0; g < ngroups; ++g)
        for(int64_t it = gstart[g]; it < gstart[g + 1]; ++it) one_item(src, dst, records + it * XO_REC);
}

int xo_rec_size(void) { return XO_REC; }
int xo_max_legs(void) { return XO_MAXL; }
