// ab200_system.cpp -- parser for the text system description (see ab200_system.h).
#include "ab200_system.h"
#include <sstream>
#include <cstdlib>
#include <cstring>

namespace ab200 {

static int opcode(const std::string& t) {
    static const char* names[] = {"CNST","VAR","ADD","SUB","MUL","DIV","NEG","POS","REC","SQR","SQRT","EXP","LOG","SIN","COS","TAN","ABS","MAX","MIN"};
    for (int i = 0; i < 19; ++i) if (t == names[i]) return i;
    return -1;
}

// A tape is a postfix expression over the nvars state variables of its location; it must leave exactly one value and
// never need more than AB_STACK slots (the evaluators index fixed-size stacks and the variable array unchecked).
static bool parse_tape(std::istringstream& in, SystemDesc& d, Func& f, int nvars, std::string& err) {
    f.begin = d.n_ops;
    std::string t;
    int depth = 0;
    while (in >> t) {
        if (t == ";") {
            if (depth != 1) { err = "tape does not evaluate to one value"; return false; }
            f.end = d.n_ops; return true;
        }
        if (d.n_ops >= AB_MAXOPS) { err = "too many tape ops"; return false; }
        TapeOp o; o.op = 0; o.arg = 0; o.lo = 0; o.hi = 0;
        if (t[0] == 'V' && t.size() > 1 && isdigit(t[1])) { o.op = OP_VAR; o.arg = atoi(t.c_str() + 1); }
        else if (t[0] == 'C' && t.size() > 1 && t != "COS") {
            size_t comma = t.find(',');
            if (comma == std::string::npos) { err = "bad constant " + t; return false; }
            o.op = OP_CNST; o.lo = strtod(t.substr(1, comma - 1).c_str(), 0); o.hi = strtod(t.substr(comma + 1).c_str(), 0);
        } else {
            int oc = opcode(t);
            if (oc < 2) { err = "bad tape token " + t; return false; }
            o.op = oc;
        }
        if (o.op == OP_VAR && (o.arg < 0 || o.arg >= nvars)) { err = "variable index out of range in " + t; return false; }
        if (o.op == OP_CNST || o.op == OP_VAR) ++depth;
        else {
            const bool binary = (o.op == OP_ADD || o.op == OP_SUB || o.op == OP_MUL || o.op == OP_DIV || o.op == OP_MAX || o.op == OP_MIN);
            if (depth < (binary ? 2 : 1)) { err = "operator " + t + " without operands"; return false; }
            if (binary) --depth;
        }
        if (depth > AB_STACK) { err = "tape needs too deep a stack"; return false; }
        d.ops[d.n_ops++] = o;
    }
    err = "unterminated tape";
    return false;
}

std::string parse_system(const std::string& text, HostSystem& out) {
    // strip comments
    std::string clean; clean.reserve(text.size());
    bool incomment = false;
    for (size_t i = 0; i < text.size(); ++i) {
        char ch = text[i];
        if (ch == '#') incomment = true;
        if (ch == '\n') incomment = false;
        if (!incomment) clean.push_back(ch);
    }
    std::istringstream in(clean);
    memset(&out.desc, 0, sizeof(SystemDesc));
    out.loc_names.clear(); out.event_names.clear();
    std::string tok, err;
    int cur_loc = -1, cur_ev = -1;
    while (in >> tok) {
        if (tok == "system") { in >> out.name; }
        else if (tok == "locations") {
            int L; in >> L;
            if (L < 1 || L > AB_MAXLOC) return "unsupported number of locations";
            out.desc.n_locs = L; out.loc_names.resize(L); out.event_names.resize(L);
        } else if (tok == "location") {
            std::string name, k1, k2; int idx, n, E;
            in >> idx >> name >> k1 >> n >> k2 >> E;
            if (idx < 0 || idx >= out.desc.n_locs || k1 != "n" || k2 != "events") return "bad location header";
            if (n < 1 || n > AB_MAXN) return "unsupported dimension";
            if (E < 0 || E > AB_MAXEV) return "too many events in a location";
            cur_loc = idx; cur_ev = -1;
            out.loc_names[idx] = name; out.event_names[idx].resize(E);
            out.desc.loc[idx].n = n; out.desc.loc[idx].n_events = E;
        } else if (tok == "dyn") {
            int i; in >> i;
            if (cur_loc < 0 || i < 0 || i >= out.desc.loc[cur_loc].n) return "bad dyn index";
            if (!parse_tape(in, out.desc, out.desc.loc[cur_loc].dyn[i], out.desc.loc[cur_loc].n, err)) return err;
        } else if (tok == "event") {
            int k, target, id; std::string name, kk, kind, kt, ki;
            in >> k >> name >> kk >> kind >> kt >> target >> ki >> id;
            if (cur_loc < 0 || k < 0 || k >= out.desc.loc[cur_loc].n_events || kk != "kind" || kt != "target" || ki != "id") return "bad event header";
            EventDesc& e = out.desc.loc[cur_loc].ev[k];
            e.kind = kind == "invariant" ? EV_INVARIANT : kind == "urgent" ? EV_URGENT : kind == "permissive" ? EV_PERMISSIVE : -1;
            if (e.kind < 0) return "bad event kind " + kind;
            e.target = target; e.id = id;
            out.event_names[cur_loc][k] = name;
            cur_ev = k;
        } else if (tok == "guard") {
            if (cur_loc < 0 || cur_ev < 0) return "guard outside event";
            if (!parse_tape(in, out.desc, out.desc.loc[cur_loc].ev[cur_ev].guard, out.desc.loc[cur_loc].n, err)) return err;
        } else if (tok == "lie") {
            if (cur_loc < 0 || cur_ev < 0) return "lie outside event";
            if (!parse_tape(in, out.desc, out.desc.loc[cur_loc].ev[cur_ev].lie, out.desc.loc[cur_loc].n, err)) return err;
        } else if (tok == "reset") {
            int i; in >> i;
            if (cur_loc < 0 || cur_ev < 0 || i < 0 || i >= AB_MAXN) return "reset outside event";
            if (!parse_tape(in, out.desc, out.desc.loc[cur_loc].ev[cur_ev].reset[i], out.desc.loc[cur_loc].n, err)) return err;
        } else {
            return "unknown token '" + tok + "'";
        }
    }
    if (out.desc.n_locs == 0) return "no locations";
    // backward dynamics: every component followed by "* constant(-1)", which is how the reference builds it
    // (VectorFunction*ScalarFunction::constant(n,-1): f(x) times the model/interval of the constant polynomial -1)
    for (int l = 0; l < out.desc.n_locs; ++l) {
        LocDesc& L = out.desc.loc[l];
        for (int i = 0; i < L.n; ++i) {
            const int len = L.dyn[i].end - L.dyn[i].begin;
            if (len <= 0) continue;
            if (out.desc.n_ops + len + 2 > AB_MAXOPS) return "too many tape ops";
            L.dyn_back[i].begin = out.desc.n_ops;
            for (int k = 0; k < len; ++k) out.desc.ops[out.desc.n_ops++] = out.desc.ops[L.dyn[i].begin + k];
            TapeOp c; c.op = OP_CNST; c.arg = 0; c.lo = -1.0; c.hi = -1.0; out.desc.ops[out.desc.n_ops++] = c;
            TapeOp m; m.op = OP_MUL; m.arg = 0; m.lo = 0; m.hi = 0; out.desc.ops[out.desc.n_ops++] = m;
            L.dyn_back[i].end = out.desc.n_ops;
        }
    }
    for (int l = 0; l < out.desc.n_locs; ++l) {
        const LocDesc& L = out.desc.loc[l];
        if (L.n == 0) return "location without a header";
        for (int i = 0; i < L.n; ++i) if (L.dyn[i].end <= L.dyn[i].begin) return "missing dynamics in location " + out.loc_names[l];
        for (int e = 0; e < L.n_events; ++e) {
            const EventDesc& E = L.ev[e];
            if (E.guard.end <= E.guard.begin) return "event without a guard in location " + out.loc_names[l];
            if (E.target < 0 || E.target >= out.desc.n_locs) return "event target out of range in location " + out.loc_names[l];
            if (E.kind != EV_INVARIANT)
                for (int i = 0; i < out.desc.loc[E.target].n; ++i) if (E.reset[i].end <= E.reset[i].begin) return "transition without a complete reset in location " + out.loc_names[l];
        }
    }
    // DiscreteEvent orders by its string id and the evolver's finishing event is DiscreteEvent(-2) == "e-2"
    // (include/discrete_event.h:39-48, src/hybrid_evolver-image.cpp:86-88)
    // The evolver resolves ties between blocking events and orders the successors of a step by event INDEX, where the
    // reference iterates std::map / std::set<DiscreteEvent>, i.e. by name: the events of a location must come sorted by name.
    for (int l = 0; l < out.desc.n_locs; ++l)
        for (int e = 1; e < out.desc.loc[l].n_events; ++e)
            if (!(out.event_names[l][e - 1] < out.event_names[l][e]))
                return "events of location " + out.loc_names[l] + " are not in ascending name order (" + out.event_names[l][e - 1] + ", " + out.event_names[l][e] + ")";
    for (int l = 0; l < out.desc.n_locs; ++l) {
        int r = 0;
        for (int e = 0; e < out.desc.loc[l].n_events; ++e) if (out.event_names[l][e] < std::string("e-2")) ++r;
        out.desc.loc[l].finishing_rank = r;
    }
    return "";
}


static double parse_bound(const std::string& t) {
    if (t == "inf" || t == "+inf") return 1.0 / 0.0;
    if (t == "-inf") return -1.0 / 0.0;
    return strtod(t.c_str(), 0);
}

std::string parse_constraints(const std::string& text, const HostSystem& sys, HostConstraints& out) {
    std::string clean; clean.reserve(text.size());
    bool incomment = false;
    for (size_t i = 0; i < text.size(); ++i) {
        char ch = text[i];
        if (ch == '#') incomment = true;
        if (ch == '\n') incomment = false;
        if (!incomment) clean.push_back(ch);
    }
    std::istringstream in(clean);
    out.sets.clear(); out.ops.clear();
    // the tapes are parsed into a scratch SystemDesc (parse_tape appends to its op array) and copied out
    SystemDesc* scratch = new SystemDesc();
    memset(scratch, 0, sizeof(SystemDesc));
    std::string tok, err;
    int cur = -1;
    while (in >> tok) {
        if (tok == "constraints") continue;
        if (tok == "set") {
            std::string name, k; int m;
            in >> name >> k >> m;
            if (k != "m" || m < 0 || m > AB_MAXN) { delete scratch; return "bad set header"; }
            int li = -1;
            for (size_t l = 0; l < sys.loc_names.size(); ++l) if (sys.loc_names[l] == name) li = (int)l;
            if (li < 0) { delete scratch; return "unknown location " + name; }
            ConstraintSetDesc d; memset(&d, 0, sizeof(d)); d.loc = li; d.m = m;
            out.sets.push_back(d); cur = (int)out.sets.size() - 1;
        } else if (cur < 0) { delete scratch; return "expected 'set' before " + tok; }
        else if (tok == "domain") {
            const int n = sys.desc.loc[out.sets[cur].loc].n;
            for (int i = 0; i < 2 * n; ++i) { std::string t; if (!(in >> t)) { delete scratch; return "short domain"; } out.sets[cur].domain[i] = parse_bound(t); }
        } else if (tok == "codomain") {
            for (int i = 0; i < 2 * out.sets[cur].m; ++i) { std::string t; if (!(in >> t)) { delete scratch; return "short codomain"; } out.sets[cur].codomain[i] = parse_bound(t); }
        } else if (tok == "f") {
            int i; in >> i;
            if (i < 0 || i >= out.sets[cur].m) { delete scratch; return "bad function index"; }
            if (!parse_tape(in, *scratch, out.sets[cur].f[i], sys.desc.loc[out.sets[cur].loc].n, err)) { delete scratch; return err; }
        } else { delete scratch; return "unknown token " + tok; }
    }
    for (size_t k = 0; k < out.sets.size(); ++k) for (int i = 0; i < out.sets[k].m; ++i)
        if (out.sets[k].f[i].end <= out.sets[k].f[i].begin) { delete scratch; return "missing constraint function"; }
    out.ops.assign(scratch->ops, scratch->ops + scratch->n_ops);
    delete scratch;
    return "";
}

} // namespace ab200
