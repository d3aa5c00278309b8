/*
 * mkcvm.c -- write a synthetic layered CVM (material) etree that the UNMODIFIED
 * reference psolve can mesh from.  TEST INFRASTRUCTURE ONLY.
 *
 * Links the reference's own libetree + cvm.c (compiled where they lie by
 * oracle/Makefile), so the file format is the reference's by construction:
 *   - schema "float Vp; float Vs; float density;" and 12-byte payload
 *     (quake/cvm/cvm.h:35-37 cvmpayload_t; shipped examples/simple/simple_case.e)
 *   - 14-token application metadata written by cvm_setdbctl (quake/cvm/cvm.c:51-80)
 *   - octants appended in Z-order inside one append transaction (etree/etree.h:441-497)
 *
 * usage: mkcvm <out.e> <level> <nx_east> <ny_north> <nz_depth> <east_m> <nlayers> {ztop_m vp vs rho}...
 *   Leaves are written at <level>; the region spans nx*ny*nz leaves, so the etree tick is
 *   east_m / (nx << (31-level)) (cvm_query derives it the same way, cvm.c:285).
 *   A leaf takes the material of the last layer whose ztop_m <= depth of the leaf centre.
 *   MKCVM_BASIN="e0,e1,n0,n1,zbot,vp,vs,rho" (metres, optional): leaves whose centre lies in that box
 *   (east e0..e1, north n0..n1, depth 0..zbot) take this material instead -- a laterally varying
 *   model, so that octor refines next to coarser octants sideways as well as downwards.
 */
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "etree.h"
#include "cvm.h"

static uint32_t compact3(uint64_t m)
{
    uint32_t v = 0;
    for (int b = 0; b < 21; b++) v |= (uint32_t)((m >> (3 * b)) & 1u) << b;
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 8) {
        fprintf(stderr, "usage: %s out.e level nx ny nz east_m nlayers {ztop vp vs rho}...\n", argv[0]);
        return 2;
    }
    const char *path = argv[1];
    int level = atoi(argv[2]);
    uint32_t nx = (uint32_t)atol(argv[3]), ny = (uint32_t)atol(argv[4]), nz = (uint32_t)atol(argv[5]);
    double east_m = atof(argv[6]);
    int nl = atoi(argv[7]);
    if (argc != 8 + 4 * nl || nl < 1 || level < 0 || level > 20) {
        fprintf(stderr, "mkcvm: bad arguments\n");
        return 2;
    }
    double *ztop = malloc(sizeof(double) * nl);
    cvmpayload_t *mat = malloc(sizeof(cvmpayload_t) * nl);
    for (int i = 0; i < nl; i++) {
        ztop[i]    = atof(argv[8 + 4 * i]);
        mat[i].Vp  = (float)atof(argv[9 + 4 * i]);
        mat[i].Vs  = (float)atof(argv[10 + 4 * i]);
        mat[i].rho = (float)atof(argv[11 + 4 * i]);
    }
    uint32_t leafticks = (uint32_t)1 << (31 - level);
    if ((uint64_t)nx * leafticks > 2147483648ull || (uint64_t)ny * leafticks > 2147483648ull ||
        (uint64_t)nz * leafticks > 2147483648ull) {
        fprintf(stderr, "mkcvm: region exceeds the etree address space at this level\n");
        return 2;
    }
    double leaf_m = east_m / nx;
    double basin[8];
    int have_basin = 0;
    if (getenv("MKCVM_BASIN") &&
        sscanf(getenv("MKCVM_BASIN"), "%lf,%lf,%lf,%lf,%lf,%lf,%lf,%lf", &basin[0], &basin[1], &basin[2], &basin[3],
               &basin[4], &basin[5], &basin[6], &basin[7]) == 8)
        have_basin = 1;
    cvmpayload_t bmat;
    bmat.Vp = (float)basin[5]; bmat.Vs = (float)basin[6]; bmat.rho = (float)basin[7];

    etree_t *ep = etree_open(path, O_CREAT | O_RDWR | O_TRUNC, 64, sizeof(cvmpayload_t), 3);
    if (!ep) { perror("etree_open"); return 1; }
    if (etree_registerschema(ep, "float Vp; float Vs; float density;") != 0) {
        fprintf(stderr, "mkcvm: %s\n", etree_strerror(etree_errno(ep)));
        return 1;
    }
    if (etree_beginappend(ep, 1.0) != 0) {
        fprintf(stderr, "mkcvm: %s\n", etree_strerror(etree_errno(ep)));
        return 1;
    }
    uint32_t nmax = nx > ny ? nx : ny;
    if (nz > nmax) nmax = nz;
    int bits = 0;
    while (((uint32_t)1 << bits) < nmax) bits++;
    uint64_t total = (uint64_t)1 << (3 * bits), count = 0;
    for (uint64_t m = 0; m < total; m++) {
        /* etree Z-order: x is the least significant interleaved bit (etree/code.c) */
        uint32_t kx = compact3(m), ky = compact3(m >> 1), kz = compact3(m >> 2);
        if (kx >= nx || ky >= ny || kz >= nz) continue;
        double zc = (kz + 0.5) * leaf_m;
        int li = 0;
        for (int i = 0; i < nl; i++) if (ztop[i] <= zc) li = i;
        etree_addr_t a;
        a.x = kx * leafticks; a.y = ky * leafticks; a.z = kz * leafticks;
        a.t = 0; a.level = level; a.type = ETREE_LEAF;
        const double xc = (kx + 0.5) * leaf_m, yc = (ky + 0.5) * leaf_m;      /* etree x = east, y = north */
        const int inb = have_basin && xc >= basin[0] && xc < basin[1] && yc >= basin[2] && yc < basin[3] && zc < basin[4];
        if (etree_append(ep, a, inb ? &bmat : &mat[li]) != 0) {
            fprintf(stderr, "mkcvm: append: %s\n", etree_strerror(etree_errno(ep)));
            return 1;
        }
        count++;
    }
    etree_endappend(ep);

    dbctl_t ctl;
    memset(&ctl, 0, sizeof ctl);
    ctl.create_model_name  = "Title:SYNTHETIC";
    ctl.create_author      = "Author:hercules-b200";
    ctl.create_date        = "Date:2026";
    ctl.create_field_count = "3";
    ctl.create_field_names = "Vp(float);Vs(float);density(float)";
    ctl.region_origin_latitude_deg  = 0;
    ctl.region_origin_longitude_deg = 0;
    ctl.region_length_east_m  = east_m;
    ctl.region_length_north_m = leaf_m * ny;
    ctl.region_depth_shallow_m = 0;
    ctl.region_depth_deep_m    = leaf_m * nz;
    ctl.domain_endpoint_x = nx * leafticks;
    ctl.domain_endpoint_y = ny * leafticks;
    ctl.domain_endpoint_z = nz * leafticks;
    if (cvm_setdbctl(ep, &ctl) != 0) return 1;
    etree_close(ep);
    printf("mkcvm: wrote %llu leaves at level %d to %s (leaf %.6g m)\n",
           (unsigned long long)count, level, path, leaf_m);
    return 0;
}
