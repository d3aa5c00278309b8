/*
 * io_planes_gpu.c -- the reference's io_planes.c, compiled WHERE IT LIES (#include REF_IO_PLANES_C, set by
 * integration/Makefile; no reference text is copied), plus ONE accessor for the GPU time loop.
 *
 * planes_print (io_planes.c:151-250) interpolates every plane point from the 8 nodes of its element, reading
 * mySolver->tm1 on the host.  Its point tables (thePlanes[].strip[][].nodestointerpolate) are file-static, so
 * the list of nodes a plane step reads can only be produced inside this translation unit.  With it,
 * psolve_gpu fetches exactly those rows of tm1 from the device (hgpu_fetch_nodes) instead of the whole field
 * and then calls the reference's own planes_print, unchanged: same interpolation, same strips, same files.
 */
#include REF_IO_PLANES_C

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
