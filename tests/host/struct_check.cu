// struct_check.cu -- HOST emulation of the structured-tile path of step_kernel (hgpu_kernels.cuh),
// compiled by nvcc for the CPU and run by tests/test_struct_host.py (no GPU needed).
//
// It uses the kernel's own helpers (sp_of_slot, gather_face, wht_forward, scale_modes,
// wht_inverse) and mirrors the kernel's thread mapping and pass order (two rounds of a dx = 0 and a dx = 1
// pass, a __syncwarp between the dy = 0 and dy = 1 halves of a pass) to check, for one aligned 8x8x8 cell:
//   1. the padded layout: 729 slots map to 729 different offsets inside a plane of SP_C doubles;
//   2. bank conflicts: every gather / accumulator access of a warp is conflict-free (16 lanes of a
//      half-warp hit 16 different 8-byte banks);
//   3. no read-modify-write at all: every element force is stored once, every node sum is formed by one thread;
//   4. numerics: the accumulated forces equal those of 512 single-element evaluations
//      (wht_forward, scale_modes, wht_inverse, scatter) to rounding.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>

#include "hgpu_internal.h"
#include "hgpu_kernels.cuh"

using namespace hgpu;

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { fails++; if (fails < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

struct Access { int tid, off; };

// conflict-free = within each half-warp the 16 offsets differ modulo 16
static void check_banks(const std::vector<Access> &acc, const char *what)
{
    std::map<int, std::vector<int>> byhalf;
    for (const Access &a : acc) byhalf[a.tid / 16].push_back(a.off);
    for (auto &kv : byhalf) {
        std::set<int> banks;
        for (int o : kv.second) banks.insert(((o % 16) + 16) % 16);
        CHECK(banks.size() == kv.second.size(), "%s: bank conflict in half-warp %d (%zu lanes, %zu banks)", what, kv.first,
              kv.second.size(), banks.size());
    }
}

static void check_banks_rebased(const std::vector<Access> &acc, int base, const char *what)
{
    std::vector<Access> b;
    for (const Access &a : acc) b.push_back({a.tid - base, a.off});
    check_banks(b, what);
}

int main()
{
    // ---- 1. layout -------------------------------------------------------------------------------
    {
        std::set<int> seen;
        for (int s = 0; s < STRUCT_NODES; s++) {
            const int sp = sp_of_slot(s);
            CHECK(sp >= 0 && sp < SP_C, "slot %d -> offset %d out of the plane", s, sp);
            CHECK(seen.insert(sp).second, "slot %d -> offset %d used twice", s, sp);
        }
        for (int z = 0; z < 9; z++) for (int y = 0; y < 9; y++) for (int x = 0; x < 9; x++)
            CHECK(sp_of_slot(struct_slot(x, y, z)) == z * SP_Z + y * SP_ROW + x, "slot of (%d,%d,%d)", x, y, z);
    }
    // ---- random w on the 729 nodes, padded planes -------------------------------------------------
    std::vector<double> W(SP_TOTAL, 0.0), acc(SP_TOTAL, 0.0), ref(3 * STRUCT_NODES, 0.0);
    srand(12345);
    auto rnd = []() { return (double)rand() / RAND_MAX - 0.5; };
    std::vector<double> wnode(3 * STRUCT_NODES);
    for (double &v : wnode) v = rnd();
    for (int z = 0; z < 9; z++) for (int y = 0; y < 9; y++) for (int x = 0; x < 9; x++)
        for (int c = 0; c < 3; c++) W[c * SP_C + z * SP_Z + y * SP_ROW + x] = wnode[3 * ((z * 9 + y) * 9 + x) + c];
    const double c1 = 1.7e9, c2 = 2.9e9;
    const double ca = -0.5625 * (c2 + 2.0 * c1), cc = -0.5625 * c2, cb = -0.5625 * c1;

    // ---- reference: 512 single elements -------------------------------------------------------------
    for (int z = 0; z < 8; z++) for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
        double w[3][8], t[3][8], v[3][8], f[3][8];
        for (int j = 0; j < 8; j++)
            for (int c = 0; c < 3; c++) w[c][j] = wnode[3 * (((z + (j >> 2)) * 9 + y + ((j >> 1) & 1)) * 9 + x + (j & 1)) + c];
        wht_forward(w[0], t[0]); wht_forward(w[1], t[1]); wht_forward(w[2], t[2]);
        scale_modes(t[0], t[1], t[2], ca, cc, cb, v[0], v[1], v[2]);
        wht_inverse(v[0], f[0]); wht_inverse(v[1], f[1]); wht_inverse(v[2], f[2]);
        for (int j = 0; j < 8; j++)
            for (int c = 0; c < 3; c++) ref[3 * (((z + (j >> 2)) * 9 + y + ((j >> 1) & 1)) * 9 + x + (j & 1)) + c] += f[c][j];
    }

    // ---- the kernel's schedule, 512 threads: thread = element (x, y, z) stores its 24 corner forces in
    //      F[(3 j + c) * 512 + thread]; after a barrier thread = node (x, y, z) (and far-face node tid < 217) sums
    //      the forces of its up to eight elements in a fixed order ------------------------------------------
    std::vector<double> F(SF_TOTAL, 0.0), own(3 * 512, 0.0), pub(3 * 217, 0.0);
    std::vector<Access> g;
    for (int lvl = 0; lvl < 2; lvl++) for (int k = 0; k < 4; k++) {           // one gather instruction = (level, face corner)
        for (int w = 0; w < 16; w++) {
            g.clear();
            for (int l = 0; l < 32; l++) {
                const int tid = 32 * w + l, x = (tid & 3) | ((tid >> 3) & 4), y = (tid >> 2) & 7, z = tid >> 6;
                g.push_back({l, z * SP_Z + y * SP_ROW + x + lvl * SP_Z + (k & 1) + SP_ROW * (k >> 1)});
            }
            check_banks(g, "gather");
        }
    }
    for (int tid = 0; tid < 512; tid++) {
        const int x = (tid & 3) | ((tid >> 3) & 4), y = (tid >> 2) & 7, z = tid >> 6;
        const int o = z * SP_Z + y * SP_ROW + x;
        double wx[8], wy[8], wz[8], tx[8], ty[8], tz[8], f[3][8];
        gather_face(W.data(), o, wx[0], wx[1], wx[2], wx[3]);
        gather_face(W.data() + SP_C, o, wy[0], wy[1], wy[2], wy[3]);
        gather_face(W.data() + 2 * SP_C, o, wz[0], wz[1], wz[2], wz[3]);
        gather_face(W.data(), o + SP_Z, wx[4], wx[5], wx[6], wx[7]);
        gather_face(W.data() + SP_C, o + SP_Z, wy[4], wy[5], wy[6], wy[7]);
        gather_face(W.data() + 2 * SP_C, o + SP_Z, wz[4], wz[5], wz[6], wz[7]);
        wht_forward(wx, tx); wht_forward(wy, ty); wht_forward(wz, tz);
        scale_modes(tx, ty, tz, ca, cc, cb, wx, wy, wz);
        wht_inverse(wx, f[0]); wht_inverse(wy, f[1]); wht_inverse(wz, f[2]);
        for (int j = 0; j < 8; j++) for (int c = 0; c < 3; c++) F[(3 * j + c) * 512 + tid] = f[c][j];
    }
    // node phase (barrier before): owned node (x, y, z) of thread tid -> Morton slot; far-face node tid -> partial force tid
    for (int j = 0; j < 8; j++) for (int w = 0; w < 16; w++) {                // bank check of the F gather, instruction by instruction
        g.clear();
        for (int l = 0; l < 32; l++) {
            const int tid = 32 * w + l, x = (tid & 3) | ((tid >> 3) & 4), y = (tid >> 2) & 7, z = tid >> 6;
            const int dx = j & 1, dy = (j >> 1) & 1, dz = j >> 2;
            if (x >= dx && y >= dy && z >= dz) {
                const int ex = x - dx;
                g.push_back({l, (ex & 3) | ((ex & 4) << 3) | ((y - dy) << 2) | ((z - dz) << 6)});
            }
        }
        check_banks(g, "node gather");
    }
    for (int tid = 0; tid < 512; tid++) {
        const int x = (tid & 3) | ((tid >> 3) & 4), y = (tid >> 2) & 7, z = tid >> 6;
        double T[3] = {0, 0, 0};
        for (int j = 0; j < 8; j++) {
            const int dx = j & 1, dy = (j >> 1) & 1, dz = j >> 2;
            if (x >= dx && y >= dy && z >= dz) {
                const int ex = x - dx, te = (ex & 3) | ((ex & 4) << 3) | ((y - dy) << 2) | ((z - dz) << 6);
                for (int c = 0; c < 3; c++) T[c] += F[(3 * j + c) * 512 + te];
            }
        }
        const int k = 3 * ((x & 1) | ((y & 1) << 1) | ((z & 1) << 2) | ((x & 2) << 2) | ((y & 2) << 3) | ((z & 2) << 4) |
                           ((x & 4) << 4) | ((y & 4) << 5) | ((z & 4) << 6));
        CHECK(k == 3 * struct_morton3(x, y, z), "Morton slot of node (%d,%d,%d)", x, y, z);
        for (int c = 0; c < 3; c++) own[k + c] = T[c];
    }
    for (int tid = 0; tid < 217; tid++) {
        int X, Y, Z;
        if (tid < 81)       { Z = 8; Y = tid / 9; X = tid - 9 * Y; }
        else if (tid < 153) { const int q = tid - 81; Y = 8; Z = q / 9; X = q - 9 * Z; }
        else                { const int q = tid - 153; X = 8; Z = q >> 3; Y = q & 7; }
        CHECK(struct_slot(X, Y, Z) == 512 + tid, "far-face node %d decodes to (%d,%d,%d)", tid, X, Y, Z);
        double T[3] = {0, 0, 0};
        for (int j = 0; j < 8; j++) {
            const int ex = X - (j & 1), ey = Y - ((j >> 1) & 1), ez = Z - (j >> 2);
            if (ex >= 0 && ex < 8 && ey >= 0 && ey < 8 && ez >= 0 && ez < 8) {
                const int te = (ex & 3) | ((ex & 4) << 3) | (ey << 2) | (ez << 6);
                for (int c = 0; c < 3; c++) T[c] += F[(3 * j + c) * 512 + te];
            }
        }
        for (int c = 0; c < 3; c++) pub[3 * tid + c] = T[c];
    }
    // back into the padded planes for the comparison below
    for (int z = 0; z < 9; z++) for (int y = 0; y < 9; y++) for (int x = 0; x < 9; x++) for (int c = 0; c < 3; c++) {
        const int sl = struct_slot(x, y, z);
        acc[c * SP_C + z * SP_Z + y * SP_ROW + x] = sl < 512 ? own[3 * sl + c] : pub[3 * (sl - 512) + c];
    }

    // ---- 4. numerics ---------------------------------------------------------------------------------
    double worst = 0.0, scale = 0.0;
    for (double v : ref) scale = std::fmax(scale, std::fabs(v));
    for (int z = 0; z < 9; z++) for (int y = 0; y < 9; y++) for (int x = 0; x < 9; x++)
        for (int c = 0; c < 3; c++)
            worst = std::fmax(worst, std::fabs(acc[c * SP_C + z * SP_Z + y * SP_ROW + x] - ref[3 * ((z * 9 + y) * 9 + x) + c]));
    CHECK(worst <= 1e-13 * scale, "forces differ: max abs diff %.3e, scale %.3e", worst, scale);
    // nothing outside the 729 node positions was touched
    {
        std::set<int> pos;
        for (int s = 0; s < STRUCT_NODES; s++) pos.insert(sp_of_slot(s));
        for (int c = 0; c < 3; c++) for (int o = 0; o < SP_C; o++)
            if (!pos.count(o)) CHECK(acc[c * SP_C + o] == 0.0, "padding position %d written", o);
    }
    printf("struct_check: %s (max rel diff %.2e)\n", fails ? "FAILED" : "ok", scale > 0 ? worst / scale : 0.0);
    return fails ? 1 : 0;
}
