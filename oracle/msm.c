/*
 * oracle/msm.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restatement of the reference's CPU Pippenger, msm/pippenger.hpp:
 *   get_wval            :12-29    window extraction from LE scalar bytes
 *   window_size         :31-38    wbits = lg(n) - {3,2} heuristics
 *   integrate_buckets   :40-56    running-sum  sum_i (i+1)*bucket[i]
 *   bucket / tile       :58-106   unsigned windows, bucket[w-1] += P
 *   mult_pippenger      :218-272  serial path: top window first, `window`
 *                                 doublings between rows
 *   breakdown + grid    :160-190,274-349  threaded tiling; here the same
 *                                 nx-by-ny tile grid is evaluated by pthreads
 *                                 and stitched in the same row order.
 * Scalars are little-endian byte strings of ceil(nbits/8) bytes, already out
 * of Montgomery form (the C-ABI entry points pass mont=false,
 * poc/msm-cuda/cuda/pippenger.cu:24).
 */
#include "msm.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static size_t get_wval(const unsigned char *d, size_t off, size_t bits, size_t nbytes)
{
    /* up to 25 bits starting at bit |off|; never reads past the scalar */
    size_t byte = off / 8, ret = 0;
    for (size_t i = 0; i < 4 && byte + i < nbytes; i++)
        ret |= (size_t)d[byte + i] << (8 * i);
    ret >>= off % 8;
    return bits >= 8 * sizeof(size_t) ? ret : ret & (((size_t)1 << bits) - 1);
}

size_t oracle_msm_window_size(size_t npoints)
{
    size_t lg = 0;
    while (npoints >>= 1)
        lg++;
    return lg > 12 ? lg - 3 : (lg > 4 ? lg - 2 : (lg ? 2 : 1));
}

static void integrate_buckets(const ec_curve *c, ec_jac *out, ec_xyzz *buckets, size_t cbits)
{
    size_t n = (size_t)1 << cbits;
    ec_xyzz acc, ret;
    acc = buckets[--n];
    ret = acc;
    ec_xyzz_inf(&buckets[n]);
    while (n--) {
        ec_xyzz_add(c, &acc, &buckets[n]);
        ec_xyzz_add(c, &ret, &acc);
        ec_xyzz_inf(&buckets[n]);
    }
    ec_xyzz_to_jac(c, out, &ret);
}

/* one tile: points [0,npoints) x scalar bits [bit0, bit0+wbits) */
static void tile(const ec_curve *c, ec_jac *ret, const ec_affine *points, size_t npoints,
                 const unsigned char *scalars, size_t nbytes, ec_xyzz *buckets,
                 size_t bit0, size_t wbits, size_t cbits)
{
    for (size_t i = 0; i < npoints; i++) {
        size_t w = wbits ? get_wval(scalars + i * nbytes, bit0, wbits, nbytes) : 0;
        w &= ((size_t)1 << cbits) - 1;
        if (w)
            ec_xyzz_madd(c, &buckets[w - 1], &points[i], 0);
    }
    integrate_buckets(c, ret, buckets, cbits);
}

void oracle_msm_pippenger_serial(const ec_curve *c, ec_jac *ret, const ec_affine *points,
                                 size_t npoints, const unsigned char *scalars)
{
    const size_t nbits = c->fr->nbits, nbytes = (nbits + 7) / 8;

    ec_jac_inf(ret);
    if (npoints == 0)
        return;
    if (npoints == 1) {
        ec_jac_mul(c, ret, &points[0], scalars, nbits);
        return;
    }

    size_t window = oracle_msm_window_size(npoints);
    ec_xyzz *buckets = calloc((size_t)1 << window, sizeof(ec_xyzz));
    ec_jac p;

    /* top excess bits first (may be zero wide), then full windows downwards */
    size_t wbits = nbits % window, cbits = wbits + 1, bit0 = nbits;
    while (bit0 -= wbits) {
        tile(c, &p, points, npoints, scalars, nbytes, buckets, bit0, wbits, cbits);
        ec_jac_add(c, ret, &p);
        for (size_t i = 0; i < window; i++)
            ec_jac_dbl(c, ret);
        cbits = wbits = window;
    }
    tile(c, &p, points, npoints, scalars, nbytes, buckets, 0, wbits, cbits);
    ec_jac_add(c, ret, &p);
    free(buckets);
}

/* ---- threaded tiling ------------------------------------------------- */

static size_t num_bits(size_t l)
{
    size_t b = 0;
    while (l) {
        b++;
        l >>= 1;
    }
    return b;
}

static void breakdown(size_t nbits, size_t window, size_t ncpus, size_t *nx_, size_t *ny_,
                      size_t *wnd_)
{
    size_t nx, ny, wnd;
    if (nbits > window * ncpus) {
        nx = 1;
        wnd = num_bits(ncpus / 4);
        if (window + wnd > 18) {
            wnd = window - wnd;
        } else {
            wnd = (nbits / window + ncpus - 1) / ncpus;
            if ((nbits / (window + 1) + ncpus - 1) / ncpus < wnd)
                wnd = window + 1;
            else
                wnd = window;
        }
    } else {
        nx = 2;
        wnd = window - 2;
        while ((nbits / wnd + 1) * nx < ncpus) {
            nx += 1;
            wnd = window - num_bits(3 * nx / 2);
        }
        nx -= 1;
        wnd = window - num_bits(3 * nx / 2);
    }
    ny = nbits / wnd + 1;
    wnd = nbits / ny + 1;
    *nx_ = nx;
    *ny_ = ny;
    *wnd_ = wnd;
}

typedef struct {
    size_t x, dx, y, dy;
    ec_jac p;
} tile_t;

typedef struct {
    const ec_curve *c;
    const ec_affine *points;
    const unsigned char *scalars;
    size_t nbytes, window, total;
    tile_t *grid;
    size_t *next;
} work_t;

static void *worker(void *arg)
{
    work_t *w = arg;
    ec_xyzz *buckets = calloc((size_t)1 << w->window, sizeof(ec_xyzz));
    for (;;) {
        size_t i = __atomic_fetch_add(w->next, 1, __ATOMIC_RELAXED);
        if (i >= w->total)
            break;
        tile_t *t = &w->grid[i];
        tile(w->c, &t->p, w->points + t->x, t->dx, w->scalars + t->x * w->nbytes, w->nbytes,
             buckets, t->y, t->dy, t->dy + (t->dy < w->window));
    }
    free(buckets);
    return NULL;
}

void oracle_msm_pippenger(const ec_curve *c, ec_jac *ret, const ec_affine *points,
                          size_t npoints, const unsigned char *scalars, size_t ncpus)
{
    const size_t nbits = c->fr->nbits, nbytes = (nbits + 7) / 8;

    if (ncpus < 2 || npoints < 32) {
        oracle_msm_pippenger_serial(c, ret, points, npoints, scalars);
        return;
    }

    size_t nx, ny, window;
    breakdown(nbits, oracle_msm_window_size(npoints), ncpus, &nx, &ny, &window);

    size_t total = 0, dx = npoints / nx, y = window * (ny - 1);
    tile_t *grid = calloc(nx * ny, sizeof(tile_t));
    while (total < nx) {
        grid[total].x = total * dx;
        grid[total].dx = dx;
        grid[total].y = y;
        grid[total].dy = nbits - y;
        total++;
    }
    grid[total - 1].dx = npoints - grid[total - 1].x;
    while (y) {
        y -= window;
        for (size_t i = 0; i < nx; i++, total++) {
            grid[total].x = grid[i].x;
            grid[total].dx = grid[i].dx;
            grid[total].y = y;
            grid[total].dy = window;
        }
    }

    size_t next = 0;
    work_t w = {c, points, scalars, nbytes, window, total, grid, &next};
    size_t nthreads = ncpus < total ? ncpus : total;
    pthread_t *th = calloc(nthreads, sizeof(pthread_t));
    for (size_t i = 0; i < nthreads; i++)
        pthread_create(&th[i], NULL, worker, &w);
    for (size_t i = 0; i < nthreads; i++)
        pthread_join(th[i], NULL);
    free(th);

    /* stitch rows top-down: add the nx tiles of a row, then `window` doublings */
    ec_jac_inf(ret);
    size_t row = 0;
    for (size_t r = 0; r < ny; r++) {
        for (size_t i = 0; i < nx; i++)
            ec_jac_add(c, ret, &grid[row++].p);
        if (r + 1 < ny)
            for (size_t i = 0; i < window; i++)
                ec_jac_dbl(c, ret);
    }
    free(grid);
}

void oracle_msm_naive(const ec_curve *c, ec_jac *ret, const ec_affine *points, size_t npoints,
                      const unsigned char *scalars)
{
    const size_t nbits = c->fr->nbits, nbytes = (nbits + 7) / 8;
    ec_jac t;
    ec_jac_inf(ret);
    for (size_t i = 0; i < npoints; i++) {
        ec_jac_mul(c, &t, &points[i], scalars + i * nbytes, nbits);
        ec_jac_add(c, ret, &t);
    }
}

/* ---- test-vector helpers --------------------------------------------- */

void oracle_gen_points(const ec_curve *c, ec_affine *out, size_t ndistinct)
{
    /* out[i] = (i+1)*G, batch-normalised with Montgomery's trick */
    if (ndistinct == 0)
        return;
    ec_jac *jac = malloc(ndistinct * sizeof(ec_jac));
    ff_t *prod = malloc(ndistinct * sizeof(ff_t));
    ec_affine g = {c->gx, c->gy};
    ec_jac acc, gj;
    ec_jac_from_affine(c, &gj, &g);
    acc = gj;
    for (size_t i = 0; i < ndistinct; i++) {
        jac[i] = acc;
        ec_jac_add(c, &acc, &gj);
    }
    ff_t run;
    ff_set_one(c->fp, &run);
    for (size_t i = 0; i < ndistinct; i++) {
        prod[i] = run;
        ff_mul(c->fp, &run, &run, &jac[i].Z);
    }
    ff_t inv;
    ff_inv(c->fp, &inv, &run);
    for (size_t i = ndistinct; i--;) {
        ff_t zi, zi2, zi3;
        ff_mul(c->fp, &zi, &inv, &prod[i]);
        ff_mul(c->fp, &inv, &inv, &jac[i].Z);
        ff_sqr(c->fp, &zi2, &zi);
        ff_mul(c->fp, &zi3, &zi2, &zi);
        ff_mul(c->fp, &out[i].X, &jac[i].X, &zi2);
        ff_mul(c->fp, &out[i].Y, &jac[i].Y, &zi3);
    }
    free(prod);
    free(jac);
}

/* flat-buffer entry points for ctypes: limbs are `n` 64-bit LE words per
 * field element, points are {X,Y} / {X,Y,Z} of such elements, packed */

static const ec_curve *curve_by_id(int id)
{
    switch (id) {                                 /* ids as include/sppark_b200.h (3 = G2: ec2.c) */
    case 0: return ec_bls12_381_g1();
    case 1: return ec_pallas();
    case 4: return ec_bn254_g1();
    case 5: return ec_bls12_377_g1();
    default: return ec_vesta();
    }
}

static void unpack_affine(const ec_curve *c, ec_affine *dst, const uint64_t *src, size_t n,
                          size_t stride_bytes)
{
    const int nl = c->fp->n;
    for (size_t i = 0; i < n; i++) {
        const uint64_t *s = (const uint64_t *)((const char *)src + i * stride_bytes);
        memset(&dst[i], 0, sizeof(dst[i]));
        memcpy(dst[i].X.l, s, 8 * nl);
        memcpy(dst[i].Y.l, s + nl, 8 * nl);
    }
}

int oracle_msm(int curve_id, uint64_t *out_jac, const void *points, size_t stride_bytes,
               size_t npoints, const unsigned char *scalars, int ncpus, int algo)
{
    const ec_curve *c = curve_by_id(curve_id);
    const int nl = c->fp->n;
    ec_affine *pts = malloc((npoints ? npoints : 1) * sizeof(ec_affine));
    unpack_affine(c, pts, points, npoints, stride_bytes);
    ec_jac r;
    if (algo == 2)
        oracle_msm_naive(c, &r, pts, npoints, scalars);
    else if (algo == 1)
        oracle_msm_pippenger_serial(c, &r, pts, npoints, scalars);
    else
        oracle_msm_pippenger(c, &r, pts, npoints, scalars, ncpus);
    free(pts);
    memcpy(out_jac, r.X.l, 8 * nl);
    memcpy(out_jac + nl, r.Y.l, 8 * nl);
    memcpy(out_jac + 2 * nl, r.Z.l, 8 * nl);
    return 0;
}

void oracle_points(int curve_id, uint64_t *out, size_t ndistinct)
{
    const ec_curve *c = curve_by_id(curve_id);
    const int nl = c->fp->n;
    ec_affine *pts = malloc((ndistinct ? ndistinct : 1) * sizeof(ec_affine));
    oracle_gen_points(c, pts, ndistinct);
    for (size_t i = 0; i < ndistinct; i++) {
        memcpy(out + 2 * nl * i, pts[i].X.l, 8 * nl);
        memcpy(out + 2 * nl * i + nl, pts[i].Y.l, 8 * nl);
    }
    free(pts);
}

void oracle_jac_to_affine(int curve_id, uint64_t *out_xy, const uint64_t *jac)
{
    const ec_curve *c = curve_by_id(curve_id);
    const int nl = c->fp->n;
    ec_jac p;
    ec_affine a;
    memset(&p, 0, sizeof(p));
    memcpy(p.X.l, jac, 8 * nl);
    memcpy(p.Y.l, jac + nl, 8 * nl);
    memcpy(p.Z.l, jac + 2 * nl, 8 * nl);
    ec_jac_to_affine(c, &a, &p);
    memcpy(out_xy, a.X.l, 8 * nl);
    memcpy(out_xy + nl, a.Y.l, 8 * nl);
}

int oracle_affine_on_curve(int curve_id, const uint64_t *xy)
{
    const ec_curve *c = curve_by_id(curve_id);
    const int nl = c->fp->n;
    ec_affine a;
    memset(&a, 0, sizeof(a));
    memcpy(a.X.l, xy, 8 * nl);
    memcpy(a.Y.l, xy + nl, 8 * nl);
    return ec_affine_on_curve(c, &a);
}

/* field KAT access: op 0 mul, 1 add, 2 sub, 3 to_mont, 4 from_mont, 5 inv */
static const ff_ctx *field_by_id(int id)
{
    switch (id) {
    case 0: return ff_bls12_381_fp();
    case 1: return ff_bls12_381_fr();
    case 2: return ff_pallas_fp();
    case 4: return ff_bn254_fp();
    case 5: return ff_bn254_fr();
    case 6: return ff_bls12_377_fp();
    case 7: return ff_bls12_377_fr();
    default: return ff_vesta_fp();
    }
}

void oracle_ff_op(int field_id, int op, uint64_t *r, const uint64_t *a, const uint64_t *b)
{
    const ff_ctx *f = field_by_id(field_id);
    ff_t x, y, z;
    ff_set_zero(&x);
    ff_set_zero(&y);
    memcpy(x.l, a, 8 * f->n);
    if (b)
        memcpy(y.l, b, 8 * f->n);
    switch (op) {
    case 0: ff_mul(f, &z, &x, &y); break;
    case 1: ff_add(f, &z, &x, &y); break;
    case 2: ff_sub(f, &z, &x, &y); break;
    case 3: ff_to_mont(f, &z, &x); break;
    case 4: ff_from_mont(f, &z, &x); break;
    default: ff_inv(f, &z, &x); break;
    }
    memcpy(r, z.l, 8 * f->n);
}

void oracle_ff_consts(int field_id, uint64_t *p, uint64_t *m0, uint64_t *rr, uint64_t *one)
{
    const ff_ctx *f = field_by_id(field_id);
    memcpy(p, f->p, 8 * f->n);
    *m0 = f->m0;
    memcpy(rr, f->rr, 8 * f->n);
    memcpy(one, f->one, 8 * f->n);
}
