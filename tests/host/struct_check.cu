// struct_check.cu -- HOST emulation of the structured-tile path of step_kernel (hgpu_kernels.cuh),
// compiled by nvcc for the CPU and run by tests/test_struct_host.py (no GPU needed).
//
// It uses the kernel's own helpers (sp_of_slot, struct_element, zsplit, face_inverse, wht_*,
// scale_modes) and mirrors the kernel's thread mapping and pass order (A, B, C, D with a __syncwarp
// between the dy = 0 and dy = 1 halves) to check, for one aligned 8x8x8 cell:
//   1. the padded layout: 729 slots map to 729 different offsets inside a plane of SP_C doubles;
//   2. bank conflicts: every gather / accumulator access of a warp is conflict-free (16 lanes of a
//      half-warp hit 16 different 8-byte banks);
//   3. races: inside one pass half, no two threads touch the same accumulator; inside one pass, threads
//      of DIFFERENT warps never touch the same accumulator (only __syncwarp orders the halves);
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

    // ---- the kernel's schedule, 256 threads ----------------------------------------------------------
    struct Regs { double wk[3][4], mid[3][4], fl[3][4], fm[3][4], ft[3][4]; int o0; };
    std::vector<Regs> R(256);
    // lower element + gathers
    std::vector<Access> g;
    for (int lvl = 0; lvl < 3; lvl++) for (int k = 0; k < 4; k++) {       // one gather instruction = (level, face corner)
        g.clear();
        for (int tid = 0; tid < 256; tid++) {
            const int x = (tid & 3) | ((tid >> 3) & 4), y = (tid >> 2) & 7, zq = tid >> 6;
            const int o0 = (2 * zq) * SP_Z + y * SP_ROW + x;
            g.push_back({tid, o0 + lvl * SP_Z + (k & 1) + SP_ROW * (k >> 1)});
        }
        check_banks(g, "gather");
    }
    for (int tid = 0; tid < 256; tid++) {
        const int x = (tid & 3) | ((tid >> 3) & 4), y = (tid >> 2) & 7, zq = tid >> 6;
        Regs &r = R[tid];
        r.o0 = (2 * zq) * SP_Z + y * SP_ROW + x;
        double wx[8], wy[8], wz[8], lo[3][4];
        gather_face(W.data(), r.o0, wx[0], wx[1], wx[2], wx[3]);
        gather_face(W.data() + SP_C, r.o0, wy[0], wy[1], wy[2], wy[3]);
        gather_face(W.data() + 2 * SP_C, r.o0, wz[0], wz[1], wz[2], wz[3]);
        gather_face(W.data(), r.o0 + SP_Z, wx[4], wx[5], wx[6], wx[7]);
        gather_face(W.data() + SP_C, r.o0 + SP_Z, wy[4], wy[5], wy[6], wy[7]);
        gather_face(W.data() + 2 * SP_C, r.o0 + SP_Z, wz[4], wz[5], wz[6], wz[7]);
        struct_element(wx, wy, wz, ca, cc, cb, lo, r.mid);
        for (int c = 0; c < 3; c++) face_inverse(lo[c], r.fl[c]);
        for (int j = 0; j < 4; j++) { r.wk[0][j] = wx[4 + j]; r.wk[1][j] = wy[4 + j]; r.wk[2][j] = wz[4 + j]; }
    }
    // passes: (name, per half: list of (level offset, which array, k))
    auto run_pass = [&](const char *name, int dx, bool upper) {
        std::map<int, int> warp_of;                     // offset -> warp that touched it in this pass
        for (int half = 0; half < 2; half++) {
            const int k = dx + 2 * half, d = dx + SP_ROW * half;
            std::map<int, int> lane_of;                 // offset -> tid within this half
            std::vector<Access> a1, a2;
            for (int tid = 0; tid < 256; tid++) {
                Regs &r = R[tid];
                std::vector<std::pair<int, const double (*)[4]>> tg;
                if (!upper) tg.push_back({r.o0 + d, r.fl});
                else { tg.push_back({r.o0 + SP_Z + d, r.fm}); tg.push_back({r.o0 + 2 * SP_Z + d, r.ft}); }
                for (size_t i = 0; i < tg.size(); i++) {
                    const int o = tg[i].first;
                    CHECK(lane_of.insert({o, tid}).second, "%s half %d: threads %d and %d update the same accumulator", name, half, lane_of[o], tid);
                    auto w = warp_of.find(o);
                    CHECK(w == warp_of.end() || w->second == tid / 32, "%s: warps %d and %d update the same accumulator inside one pass",
                          name, w == warp_of.end() ? -1 : w->second, tid / 32);
                    warp_of[o] = tid / 32;
                    acc_add3(acc.data(), o, tg[i].second[0][k], tg[i].second[1][k], tg[i].second[2][k]);
                    (i == 0 ? a1 : a2).push_back({tid, o});
                }
            }
            check_banks(a1, name);
            if (!a2.empty()) check_banks(a2, name);
        }
    };
    run_pass("pass A", 0, false);
    run_pass("pass B", 1, false);
    // upper element
    for (int tid = 0; tid < 256; tid++) {
        Regs &r = R[tid];
        double wx[8], wy[8], wz[8], lo[3][4], top[3][4];
        for (int j = 0; j < 4; j++) { wx[j] = r.wk[0][j]; wy[j] = r.wk[1][j]; wz[j] = r.wk[2][j]; }
        gather_face(W.data(), r.o0 + 2 * SP_Z, wx[4], wx[5], wx[6], wx[7]);
        gather_face(W.data() + SP_C, r.o0 + 2 * SP_Z, wy[4], wy[5], wy[6], wy[7]);
        gather_face(W.data() + 2 * SP_C, r.o0 + 2 * SP_Z, wz[4], wz[5], wz[6], wz[7]);
        struct_element(wx, wy, wz, ca, cc, cb, lo, top);
        for (int c = 0; c < 3; c++) {
            for (int k = 0; k < 4; k++) r.mid[c][k] += lo[c][k];
            face_inverse(r.mid[c], r.fm[c]); face_inverse(top[c], r.ft[c]);
        }
    }
    run_pass("pass C", 0, true);
    run_pass("pass D", 1, true);

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
