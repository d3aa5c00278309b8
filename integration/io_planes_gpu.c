/*
 * io_planes_gpu.c -- the reference's io_planes.c, compiled WHERE IT LIES (#include REF_IO_PLANES_C, set by
 * integration/Makefile; no reference text is copied), plus the accessors the GPU time loop needs: the node
 * list of the sparse fetch, and the point tables + strip transport of the device-interpolated planes.
 *
 * planes_print (io_planes.c:151-250) interpolates every plane point from the 8 nodes of its element, reading
 * mySolver->tm1 on the host.  Its point tables (thePlanes[].strip[][].nodestointerpolate) are file-static, so
 * the list of nodes a plane step reads can only be produced inside this translation unit.  With it,
 * psolve_gpu fetches exactly those rows of tm1 from the device (hgpu_fetch_nodes) instead of the whole field
 * and then calls the reference's own planes_print, unchanged: same interpolation, same strips, same files.
 */
#ifdef HGPU_PLANES_HOSTCHECK
#define planes_print ref_planes_print      /* the reference's entry point keeps its body under another name */
#endif
#include REF_IO_PLANES_C
#ifdef HGPU_PLANES_HOSTCHECK
#undef planes_print
#endif

/* Local node ids read by planes_print on this rank (duplicates included), 8 per plane point, in plane /
 * strip / point order.  out == NULL: only count.  Old (single I/O rank) layout only: with IO_pool_pe_count
 * != 0 the reference uses the strips differently (io_planes.c:475-650) and the caller keeps fetching whole
 * fields. */
int64_t hgpu_planes_node_list(int theNumberOfPlanes, int32_t *out)
{
    int64_t n = 0;
    if (thePlanes == NULL) return 0;
    for (int p = 0; p < theNumberOfPlanes; p++)
        for (int s = 0; s < thePlanes[p].numberofstripsthisplane; s++) {
            const int len = thePlanes[p].stripend[s] - thePlanes[p].stripstart[s] + 1;
            for (int e = 0; e < len; e++)
                for (int k = 0; k < 8; k++) {
                    if (out) out[n] = thePlanes[p].strip[s][e].nodestointerpolate[k];
                    n++;
                }
        }
    return n;
}

/* ---- planes interpolated on the device (hgpu_planes_*, include/hercules_gpu.h) ---------------------------
 * The point tables of this rank in plane / strip / point order: nodes [npoints][8] and localcoords
 * [npoints][3] (plane_strip_element_t, io_planes.c:83-88) for hgpu_planes_attach.  NULL pointers: only count. */
int64_t hgpu_planes_point_tables(int theNumberOfPlanes, int32_t *nodes, double *local)
{
    int64_t n = 0;
    if (thePlanes == NULL) return 0;
    for (int p = 0; p < theNumberOfPlanes; p++)
        for (int s = 0; s < thePlanes[p].numberofstripsthisplane; s++) {
            const int len = thePlanes[p].stripend[s] - thePlanes[p].stripstart[s] + 1;
            for (int e = 0; e < len; e++, n++) {
                if (nodes)
                    for (int k = 0; k < 8; k++) nodes[8 * n + k] = thePlanes[p].strip[s][e].nodestointerpolate[k];
                if (local)
                    for (int c = 0; c < 3; c++) local[3 * n + c] = thePlanes[p].strip[s][e].localcoords.x[c];
            }
        }
    return n;
}

/* What Old_planes_print does AFTER interpolating a strip (io_planes.c:193-247), on rows the device
 * interpolated (hgpu_planes_record, same point order as hgpu_planes_point_tables): rank 0 places its own
 * strips in the plane buffer, the other ranks send theirs with the start location appended (the reference's
 * wire format: 3 len doubles + start + 0.1, tag = plane), rank 0 receives what it did not produce itself and
 * the reference's own writer prints the plane. */
int hgpu_planes_print_rows(int32_t myID, int theNumberOfPlanes, const double *rows)
{
    for (int p = 0; p < theNumberOfPlanes; p++) {
        const int mine = thePlanes[p].numberofstripsthisplane;
        for (int s = 0; s < mine; s++) {
            const int start = thePlanes[p].stripstart[s];
            const size_t len3 = 3 * (size_t)(thePlanes[p].stripend[s] - start + 1);
            if (myID == 0) {
                memcpy(planes_output_buffer + 3 * (size_t)start, rows, len3 * sizeof(double));
            } else {
                memcpy(planes_stripMPISendBuffer, rows, len3 * sizeof(double));
                planes_stripMPISendBuffer[len3] = (double)(start + 0.1);
                MPI_Send(planes_stripMPISendBuffer, (int)len3 + 1, MPI_DOUBLE, 0, p, comm_solver);
            }
            rows += len3;
        }
        if (myID == 0)
            for (int left = thePlanes[p].globalnumberofstripsthisplane - mine; left > 0; left--) {
                MPI_Status st;
                int got;
                MPI_Recv(planes_stripMPIRecvBuffer, planes_GlobalLargestStripCount * 3 + 1, MPI_DOUBLE, MPI_ANY_SOURCE,
                         p, comm_solver, &st);
                MPI_Get_count(&st, MPI_DOUBLE, &got);
                memcpy(planes_output_buffer + 3 * (size_t)(int)planes_stripMPIRecvBuffer[got - 1], planes_stripMPIRecvBuffer,
                       (size_t)(got - 1) * sizeof(double));
            }
        Old_print_plane_displacements(myID, p);
        MPI_Barrier(comm_solver);
    }
    return 1;
}

#ifdef HGPU_PLANES_HOSTCHECK
/* CPU check of the two functions above (integration/_bin/psolve_planes_hostcheck = the unmodified reference
 * psolve.o + this file; tests/test_planes_host.py): planes_print as psolve_gpu performs it with
 * PSOLVE_GPU_DEVICE_PLANES=1, the device kernel replaced by the same arithmetic on the host
 * (plane_kernel, hercules_b200/csrc/hgpu_kernels.cuh: one rounded multiply / add at a time, -ffp-contract=off).
 * The plane files must be byte-identical to the reference's on any number of ranks. */
int planes_print(int32_t myID, int IO_pool_pe_count, int theNumberOfPlanes, mysolver_t *mySolver)
{
    static int64_t npts = -1;
    static int32_t *nd;
    static double *lc, *rows;
    if (IO_pool_pe_count) return ref_planes_print(myID, IO_pool_pe_count, theNumberOfPlanes, mySolver);
    if (npts < 0) {
        npts = hgpu_planes_point_tables(theNumberOfPlanes, NULL, NULL);
        nd = malloc(sizeof(int32_t) * 8 * (size_t)(npts + 1));
        lc = malloc(sizeof(double) * 3 * (size_t)(npts + 1));
        rows = malloc(sizeof(double) * 3 * (size_t)(npts + 1));
        hgpu_planes_point_tables(theNumberOfPlanes, nd, lc);
    }
    for (int64_t p = 0; p < npts; p++) {
        double d[3] = {0.0, 0.0, 0.0};
        for (int i = 0; i < 8; i++) {
            const double sx = (i & 1) ? 1.0 : -1.0, sy = (i & 2) ? 1.0 : -1.0, sz = (i & 4) ? 1.0 : -1.0;
            const double phi = (1.0 + sx * lc[3 * p]) * (1.0 + sy * lc[3 * p + 1]) * (1.0 + sz * lc[3 * p + 2]) * 0.125;
            for (int c = 0; c < 3; c++) d[c] = d[c] + phi * mySolver->tm1[nd[8 * p + i]].f[c];
        }
        for (int c = 0; c < 3; c++) rows[3 * p + c] = d[c];
    }
    return hgpu_planes_print_rows(myID, theNumberOfPlanes, rows);
}
#endif
