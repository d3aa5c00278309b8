// struct_check.cu -- HOST emulation of the structured-tile path of step_kernel (hgpu_kernels.cuh),
// compiled by nvcc for the CPU and run by tests/test_struct_host.py (no GPU needed).
//
// It uses the kernel's own helpers (sp_of_slot, gather_face, acc_add3, wht_forward, scale_modes,
// wht_inverse) and mirrors the kernel's thread mapping and pass order (two rounds of a dx = 0 and a dx = 1
// pass, a __syncwarp between the dy = 0 and dy = 1 halves of a pass) to check, for one aligned 8x8x8 cell:
//   1. the padded layout: 729 slots map to 729 different offsets inside a plane of SP_C doubles;
//   2. bank conflicts: every gather / accumulator access of a warp is conflict-free (16 lanes of a
//      half-warp hit 16 different 8-byte banks);
//   3. races: inside one ROUND (the code between two barriers) every accumulator address is updated by
//      exactly one thread of the CTA (the shares of neighbouring elements are combined with shuffles first);
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
    std::vector<double> W(SP_TOTAL, 0.0), acc(SP_TOTAL, 0.0), accx(SX4_TOTAL, 0.0), ref(3 * STRUCT_NODES, 0.0);
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

    // ---- the kernel's schedule, 256 threads: two rounds (element z = 2 zq, then 2 zq + 1), each with a
    //      dx = 0 pass and a dx = 1 pass (barrier between them), each pass with a dy = 0 and a dy = 1 half
    //      (__syncwarp between them), each half touching the element's lower and upper level ------------
    struct Regs { double f[3][8]; int o; };
    std::vector<Regs> R(256);
    std::vector<Access> g;
    for (int r = 0; r < 2; r++) {
        for (int lvl = 0; lvl < 2; lvl++) for (int k = 0; k < 4; k++) {       // one gather instruction = (level, face corner)
            g.clear();
            for (int tid = 0; tid < 256; tid++) {
                const int x = (tid & 3) | ((tid >> 3) & 4), y = (tid >> 2) & 7, zq = tid >> 6;
                const int o = (2 * zq + r) * SP_Z + y * SP_ROW + x;
                g.push_back({tid, o + lvl * SP_Z + (k & 1) + SP_ROW * (k >> 1)});
            }
            check_banks(g, "gather");
        }
        for (int tid = 0; tid < 256; tid++) {
            const int x = (tid & 3) | ((tid >> 3) & 4), y = (tid >> 2) & 7, zq = tid >> 6;
            Regs &q = R[tid];
            q.o = (2 * zq + r) * SP_Z + y * SP_ROW + x;
            double wx[8], wy[8], wz[8], tx[8], ty[8], tz[8];
            gather_face(W.data(), q.o, wx[0], wx[1], wx[2], wx[3]);
            gather_face(W.data() + SP_C, q.o, wy[0], wy[1], wy[2], wy[3]);
            gather_face(W.data() + 2 * SP_C, q.o, wz[0], wz[1], wz[2], wz[3]);
            gather_face(W.data(), q.o + SP_Z, wx[4], wx[5], wx[6], wx[7]);
            gather_face(W.data() + SP_C, q.o + SP_Z, wy[4], wy[5], wy[6], wy[7]);
            gather_face(W.data() + 2 * SP_C, q.o + SP_Z, wz[4], wz[5], wz[6], wz[7]);
            wht_forward(wx, tx); wht_forward(wy, ty); wht_forward(wz, tz);
            scale_modes(tx, ty, tz, ca, cc, cb, wx, wy, wz);
            wht_inverse(wx, q.f[0]); wht_inverse(wy, q.f[1]); wht_inverse(wz, q.f[2]);
        }
        // one ROUND = the code between two barriers.  The accumulator update is shuffle-combined: per level and
        // component the lanes exchange shares with __shfl_up (by 4 = y - 1, by 1 = x - 1) and every address is
        // updated by exactly ONE thread of the CTA in the round -- checked here over all 256 threads.
        std::map<long, int> owner;                          // address -> thread that updated it in this round
        auto upd = [&](int tid, bool side, int o, int c, double v, std::vector<Access> *rec) {
            const long key = (side ? 1000000L : 0L) + o + (side ? c * SX4_C : c * SP_C);
            CHECK(owner.insert({key, tid}).second, "round %d: threads %d and %d update the same accumulator", r, owner[key], tid);
            if (side) accx[o + c * SX4_C] += v; else { acc[o + c * SP_C] += v; if (rec) rec->push_back({tid, o}); }
        };
        for (int w = 0; w < 8; w++) for (int dz = 0; dz < 2; dz++) for (int c = 0; c < 3; c++) {
            double F0[32], F1[32], F2[32], F3[32], u3[32], u2[32], a[32], ua[32], u3x[32];
            for (int l = 0; l < 32; l++) { const Regs &q = R[32 * w + l]; F0[l] = q.f[c][4 * dz]; F1[l] = q.f[c][4 * dz + 1]; F2[l] = q.f[c][4 * dz + 2]; F3[l] = q.f[c][4 * dz + 3]; }
            for (int l = 0; l < 32; l++) { u3[l] = l >= 4 ? F3[l - 4] : F3[l]; u2[l] = l >= 4 ? F2[l - 4] : F2[l]; u3x[l] = l >= 1 ? F3[l - 1] : F3[l]; }
            for (int l = 0; l < 32; l++) a[l] = ((l >> 2) & 7) != 0 ? F1[l] + u3[l] : F1[l];
            for (int l = 0; l < 32; l++) ua[l] = l >= 1 ? a[l - 1] : a[l];
            std::vector<Access> a_own, a_x, a_y;
            for (int l = 0; l < 32; l++) {
                const int tid = 32 * w + l;
                const int x = (tid & 3) | ((tid >> 3) & 4), y = (tid >> 2) & 7, zq = tid >> 6;
                const bool xlo = (tid & 3) != 0, xhi = (tid & 3) == 3, ylo = y != 0, yhi = y == 7;
                const int ol = R[tid].o + dz * SP_Z, lx = (2 * zq + r + dz) * 9 + y;
                double own = ylo ? F0[l] + u2[l] : F0[l];
                if (xlo) own += ua[l];
                upd(tid, false, ol, c, own, &a_own);
                if (xhi) { if (x == 3) upd(tid, true, lx, c, a[l], nullptr); else upd(tid, false, ol + 1, c, a[l], &a_x); }
                if (yhi) upd(tid, false, ol + SP_ROW, c, xlo ? F2[l] + u3x[l] : F2[l], &a_y);
                if (xhi && yhi) { if (x == 3) upd(tid, true, lx + 1, c, F3[l], nullptr); else upd(tid, false, ol + SP_ROW + 1, c, F3[l], nullptr); }
            }
            check_banks_rebased(a_own, 32 * w, "own update"); check_banks_rebased(a_x, 32 * w, "+x edge update"); check_banks_rebased(a_y, 32 * w, "+y edge update");
        }
    }
    // drain: the side array belongs to the x = 4 column
    for (int z = 0; z < 9; z++) for (int y = 0; y < 9; y++) for (int c = 0; c < 3; c++)
        acc[c * SP_C + z * SP_Z + y * SP_ROW + 4] += accx[c * SX4_C + z * 9 + y];

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
