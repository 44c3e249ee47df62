// tests/tools/hostkat.cpp -- test infrastructure only (never shipped, never imported by the product path).
// Replays the known-answer vectors of the primitives (tests/golden/kat_primitives.txt, produced by the reference library
// through oracle/harness/ref_kat.cpp) through the HOST build of the kernel headers: the cooperative primitives run under the
// fiber warp emulator (tests/tools/warp_emu.h, 32 lanes) with fenv rounding, through the same dispatch the device kernels use
// (release-1.0_b200/csrc/ab200_kat.cuh). Bit-exact comparison incl. the structure of the models.
//   hostkat <kat_primitives.txt> [lanes]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include <map>
#define AB_WARP_EMU 1
#include "warp_emu.h"
#include "../../release-1.0_b200/csrc/ab200_kat.cuh"

using namespace ab200;

struct Model { std::vector<std::vector<int> > ex; std::vector<double> cf; double e; };
struct Case { std::string op; int as, k, flag; double s, il, iu; Model x, y, r; bool have_v; double vl, vu; };

static const char* TM_OPS[] = {"add", "sub", "mul", "addf", "subf", "mulf", "divf", "addi", "subi", "muli", "divi", "neg", "intersection", "split",
                               "antiderivative", "rec", "sqrt", "abs", "max", "range", "exp", "log", "sin", "cos"};
static const char* IV_OPS[] = {"add", "sub", "mul", "div", "addf", "subf", "mulf", "divf", "fsub", "fdiv", "sqr", "sqrt", "pow", "rec", "abs", "neg",
                               "max", "min", "exp", "log", "sin", "cos"};
static int code(const char** names, int n, const std::string& s) { for (int i = 0; i < n; ++i) if (s == names[i]) return i; return -1; }

static TM load(Ctx& c, int as, const Model& m) {
    TM t = tm_begin(c, as, acc_default());
    for (size_t i = 0; i < m.cf.size(); ++i) { Key k = key_zero(); for (int j = 0; j < as; ++j) if (m.ex[i][j]) key_inc(k, j, m.ex[i][j]); tm_append(c, t, k, m.cf[i]); }
    t.e = m.e;
    return t;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: hostkat <kat_primitives.txt> [lanes]\n"); return 2; }
    const int lanes = argc > 2 ? atoi(argv[2]) : 32;
    std::ifstream f(argv[1]);
    std::string line;
    std::vector<Case> cases; Model* cur = 0;
    int iv_total = 0, iv_bad = 0;
    while (std::getline(f, line)) {
        std::istringstream in(line); std::string t; in >> t;
        if (t == "tm") {
            Case c; std::string k, v; in >> c.op; c.have_v = false; c.vl = c.vu = 0;
            in >> k >> c.as >> k >> c.k >> k >> c.flag >> k >> v; c.s = strtod(v.c_str(), 0); in >> k >> v; c.il = strtod(v.c_str(), 0); in >> k >> v; c.iu = strtod(v.c_str(), 0);
            cases.push_back(c); cur = 0;
        } else if (t == "x" || t == "y" || t == "r") {
            Model& m = t == "x" ? cases.back().x : (t == "y" ? cases.back().y : cases.back().r);
            int n; std::string v; in >> n >> v; m.e = strtod(v.c_str(), 0); m.ex.clear(); m.cf.clear(); cur = &m;
        } else if (t == "t" && cur) {
            std::vector<std::string> tok; std::string v; while (in >> v) tok.push_back(v);
            std::vector<int> ex; for (size_t i = 0; i + 1 < tok.size(); ++i) ex.push_back(atoi(tok[i].c_str()));
            cur->ex.push_back(ex); cur->cf.push_back(strtod(tok.back().c_str(), 0));
        } else if (t == "v") {
            std::string a, b; in >> a >> b; cases.back().have_v = true; cases.back().vl = strtod(a.c_str(), 0); cases.back().vu = strtod(b.c_str(), 0);
        } else if (t == "iv") {
            std::string op, k, v[6]; int m;
            in >> op >> k >> v[0] >> v[1] >> k >> v[2] >> v[3] >> k >> m >> k >> v[4] >> v[5];
            double d[6]; for (int i = 0; i < 6; ++i) d[i] = strtod(v[i].c_str(), 0);
            int st = ST_OK;
            fesetround(FE_TONEAREST);
            const Ivl r = kat_interval_op(code(IV_OPS, 22, op), m, ivl(d[0], d[1]), ivl(d[2], d[3]), &st);
            ++iv_total;
            if (st != ST_OK || memcmp(&r.l, &d[4], 8) || memcmp(&r.u, &d[5], 8)) { ++iv_bad; if (iv_bad <= 10) printf("interval %s [%a,%a] [%a,%a] m=%d: [%a,%a] vs [%a,%a] status %d\n", op.c_str(), d[0], d[1], d[2], d[3], m, r.l, r.u, d[4], d[5], st); }
        }
    }
    static std::vector<uint64_t> arena(1 << 22), smem(AB_SM_WORDS);
    int bad = 0;
    for (size_t ci = 0; ci < cases.size(); ++ci) {
        const Case& k = cases[ci];
        const int op = code(TM_OPS, 24, k.op);
        int status = 0; Ivl rng = ivl(0, 0); std::vector<uint64_t> words; double err = 0; int n = 0;
        wemu::run(lanes, [&](int lane) {
            Ctx c; ctx_init(c, arena.data(), (int)arena.size(), smem.data(), binom_host()); ctx_stage_tables(c);
            TM x = load(c, k.as, k.x); wsync(); TM y = load(c, k.as, k.y); wsync();
            Ivl rg = ivl(0, 0);
            TM r = kat_tm_op(c, op, x, y, k.k, k.flag, k.s, ivl(k.il, k.iu), &rg);
            wsync();
            if (lane == 0) { status = c.status; rng = rg; n = r.n; err = r.e; words.assign(c.A.base + r.off, c.A.base + r.off + r.n * AB_TW); }
        });
        bool ok = status == ST_OK;
        if (ok && k.op == "range") ok = k.have_v && !memcmp(&rng.l, &k.vl, 8) && !memcmp(&rng.u, &k.vu, 8);
        else if (ok) {
            ok = n == (int)k.r.cf.size() && !memcmp(&err, &k.r.e, 8);
            for (int t = 0; ok && t < n; ++t) {
                Key key = key_zero(); for (int j = 0; j < k.as; ++j) if (k.r.ex[t][j]) key_inc(key, j, k.r.ex[t][j]);
                if (words[t * AB_TW] != key.w[0] || memcmp(&words[t * AB_TW + AB_KW], &k.r.cf[t], 8)) ok = false;
            }
        }
        if (!ok) { ++bad; if (bad <= 10) printf("tm #%zu %s as=%d: status %d terms %d vs %zu error %a vs %a\n", ci, k.op.c_str(), k.as, status, n, k.r.cf.size(), err, k.r.e); }
    }
    printf("hostkat: %zu TaylorModel cases, %d mismatching; %d Interval cases, %d mismatching (lanes %d)\n", cases.size(), bad, iv_total, iv_bad, lanes);
    return (bad || iv_bad) ? 1 : 0;
}
