// Oracle tooling only: a small reduced-ordered BDD package exposing the subset of the BuDDy 2.4 C++ API
// that the reference's src/bdd_set.cpp uses (BuDDy itself is absent from this image). Variable order is the
// identity (level == variable index), as in BuDDy without reordering. Nodes are never freed.
// ROBDDs are canonical, so every set operation yields the same Boolean function and therefore the same
// node structure (var/low/high) as BuDDy would.
#ifndef ORACLE_SHIM_MINIBUDDY_BDD_H
#define ORACLE_SHIM_MINIBUDDY_BDD_H
#include <vector>
#include <unordered_map>
#include <iostream>
#include <cstdlib>
#include <cstdint>
#include <algorithm>

struct bddGbcStat { int nodes; int freenodes; long time; long sumtime; int num; };
typedef void (*bddgbchandler)(int, bddGbcStat*);
typedef void (*bdderrhandler)(int);

enum { bddop_and=0, bddop_xor=1, bddop_or=2, bddop_nand=3, bddop_nor=4, bddop_imp=5, bddop_biimp=6,
       bddop_diff=7, bddop_less=8, bddop_invimp=9 };

namespace minibuddy {
struct Node { int var; int low; int high; };
struct Manager {
    bool running; int varnum;
    std::vector<Node> nodes;                          // 0 = false, 1 = true
    std::unordered_map<uint64_t,std::vector<int> > unique;
    std::unordered_map<uint64_t,int> cache[10];
    Manager() : running(false), varnum(0) { }
    static Manager& get() { static Manager m; return m; }
    void init() { nodes.clear(); unique.clear(); for(int i=0;i<10;++i) cache[i].clear();
        Node f={1<<30,0,0}; Node t={1<<30,1,1}; nodes.push_back(f); nodes.push_back(t); running=true; }
    int mk(int var, int low, int high) {
        if(low==high) return low;
        uint64_t h=(uint64_t(uint32_t(var))*0x9E3779B97F4A7C15ull) ^ (uint64_t(uint32_t(low))<<32 | uint32_t(high));
        std::vector<int>& bucket=unique[h];
        for(size_t k=0;k!=bucket.size();++k) { const Node& n=nodes[bucket[k]]; if(n.var==var && n.low==low && n.high==high) return bucket[k]; }
        Node n={var,low,high}; nodes.push_back(n); int id=int(nodes.size())-1; bucket.push_back(id); return id; }
    static bool opval(int op, bool a, bool b) {
        switch(op) { case bddop_and: return a&&b; case bddop_xor: return a!=b; case bddop_or: return a||b;
            case bddop_nand: return !(a&&b); case bddop_nor: return !(a||b); case bddop_imp: return !a||b;
            case bddop_biimp: return a==b; case bddop_diff: return a&&!b; case bddop_less: return !a&&b;
            default: return a||!b; } }
    int apply(int a, int b, int op) {
        if(a<2 && b<2) return opval(op,a==1,b==1)?1:0;
        switch(op) {
            case bddop_and: if(a==0||b==0) return 0; if(a==1) return b; if(b==1) return a; if(a==b) return a; break;
            case bddop_or: if(a==1||b==1) return 1; if(a==0) return b; if(b==0) return a; if(a==b) return a; break;
            case bddop_diff: if(a==0||b==1) return 0; if(b==0) return a; if(a==b) return 0; break;
            default: break; }
        uint64_t key=(uint64_t(uint32_t(a))<<32)|uint32_t(b);
        std::unordered_map<uint64_t,int>::iterator it=cache[op].find(key);
        if(it!=cache[op].end()) return it->second;
        int va=nodes[a].var, vb=nodes[b].var; int v=std::min(va,vb);
        int al=(va==v)?nodes[a].low:a, ah=(va==v)?nodes[a].high:a;
        int bl=(vb==v)?nodes[b].low:b, bh=(vb==v)?nodes[b].high:b;
        int l=apply(al,bl,op); int h=apply(ah,bh,op);
        int r=mk(v,l,h); cache[op][key]=r; return r; }
    // quantify: vars is a cube (conjunction of positive literals); forall -> and, exists -> or
    int quant(int a, int vars, int qop, std::unordered_map<uint64_t,int>& memo) {
        if(a<2 || vars<2) return a;
        while(vars>=2 && nodes[vars].var<nodes[a].var) vars=nodes[vars].high;
        if(vars<2) return a;
        uint64_t key=(uint64_t(uint32_t(a))<<32)|uint32_t(vars);
        std::unordered_map<uint64_t,int>::iterator it=memo.find(key); if(it!=memo.end()) return it->second;
        int l=quant(nodes[a].low,vars,qop,memo), h=quant(nodes[a].high,vars,qop,memo);
        int r=(nodes[vars].var==nodes[a].var) ? apply(l,h,qop) : mk(nodes[a].var,l,h);
        memo[key]=r; return r; }
    int ite(int f, int g, int h) { return apply(apply(f,g,bddop_and),apply(apply(f,1,bddop_xor),h,bddop_and),bddop_or); }
    int replace(int a, const std::vector<int>& map, std::unordered_map<int,int>& memo) {
        if(a<2) return a;
        std::unordered_map<int,int>::iterator it=memo.find(a); if(it!=memo.end()) return it->second;
        int l=replace(nodes[a].low,map,memo), h=replace(nodes[a].high,map,memo);
        int v=nodes[a].var; int nv=(v<int(map.size()) && map[v]>=0)?map[v]:v;
        // general (order-changing) renaming handled through ite on the new variable
        int r=ite(mk(nv,0,1),h,l); memo[a]=r; return r; }
    void support(int a, std::vector<bool>& seen, std::vector<bool>& vars) {
        if(a<2 || seen[a]) return; seen[a]=true; vars[nodes[a].var]=true;
        support(nodes[a].low,seen,vars); support(nodes[a].high,seen,vars); }
};
} // namespace minibuddy

class bdd {
    int _id;
  public:
    bdd() : _id(0) { }
    explicit bdd(int id) : _id(id) { }
    int id() const { return _id; }
    bool operator==(const bdd& o) const { return _id==o._id; }
    bool operator!=(const bdd& o) const { return _id!=o._id; }
    bdd operator&(const bdd& o) const { return bdd(minibuddy::Manager::get().apply(_id,o._id,bddop_and)); }
    bdd operator|(const bdd& o) const { return bdd(minibuddy::Manager::get().apply(_id,o._id,bddop_or)); }
    bdd operator^(const bdd& o) const { return bdd(minibuddy::Manager::get().apply(_id,o._id,bddop_xor)); }
    bdd operator-(const bdd& o) const { return bdd(minibuddy::Manager::get().apply(_id,o._id,bddop_diff)); }
    bdd operator!() const { return bdd(minibuddy::Manager::get().apply(_id,1,bddop_xor)); }
    bdd& operator&=(const bdd& o) { *this=*this&o; return *this; }
    bdd& operator|=(const bdd& o) { *this=*this|o; return *this; }
};

static const bdd bddfalse(0);
static const bdd bddtrue(1);

struct bddPair { std::vector<int> map; };

inline int bdd_isrunning() { return minibuddy::Manager::get().running ? 1 : 0; }
inline int bdd_init(int, int) { minibuddy::Manager::get().init(); return 0; }
inline int bdd_setvarnum(int n) { minibuddy::Manager& m=minibuddy::Manager::get(); if(n>m.varnum) m.varnum=n; return 0; }
inline int bdd_varnum() { return minibuddy::Manager::get().varnum; }
inline bddgbchandler bdd_gbc_hook(bddgbchandler h) { return h; }
inline bdderrhandler bdd_error_hook(bdderrhandler h) { return h; }
inline const char* bdd_errstring(int) { return "minibuddy error"; }
inline bdd bdd_true() { return bddtrue; }
inline bdd bdd_false() { return bddfalse; }
inline bdd bdd_ithvar(int v) { return bdd(minibuddy::Manager::get().mk(v,0,1)); }
inline bdd bdd_nithvar(int v) { return bdd(minibuddy::Manager::get().mk(v,1,0)); }
inline int bdd_var(const bdd& b) { return minibuddy::Manager::get().nodes[b.id()].var; }
inline bdd bdd_low(const bdd& b) { return bdd(minibuddy::Manager::get().nodes[b.id()].low); }
inline bdd bdd_high(const bdd& b) { return bdd(minibuddy::Manager::get().nodes[b.id()].high); }
inline int bdd_var2level(int v) { return v; }
inline int bdd_level2var(int l) { return l; }
inline bdd bdd_apply(const bdd& a, const bdd& b, int op) { return bdd(minibuddy::Manager::get().apply(a.id(),b.id(),op)); }
inline bdd bdd_and(const bdd& a, const bdd& b) { return bdd_apply(a,b,bddop_and); }
inline bdd bdd_or(const bdd& a, const bdd& b) { return bdd_apply(a,b,bddop_or); }
inline bdd bdd_xor(const bdd& a, const bdd& b) { return bdd_apply(a,b,bddop_xor); }
inline bdd bdd_imp(const bdd& a, const bdd& b) { return bdd_apply(a,b,bddop_imp); }
inline bdd bdd_biimp(const bdd& a, const bdd& b) { return bdd_apply(a,b,bddop_biimp); }
inline bdd bdd_not(const bdd& a) { return !a; }
inline bdd bdd_forall(const bdd& a, const bdd& vars) {
    std::unordered_map<uint64_t,int> memo; return bdd(minibuddy::Manager::get().quant(a.id(),vars.id(),bddop_and,memo)); }
inline bdd bdd_exist(const bdd& a, const bdd& vars) {
    std::unordered_map<uint64_t,int> memo; return bdd(minibuddy::Manager::get().quant(a.id(),vars.id(),bddop_or,memo)); }
inline bdd bdd_appex(const bdd& a, const bdd& b, int op, const bdd& vars) { return bdd_exist(bdd_apply(a,b,op),vars); }
inline bdd bdd_support(const bdd& a) {
    minibuddy::Manager& m=minibuddy::Manager::get();
    std::vector<bool> seen(m.nodes.size(),false); int maxv=0;
    for(size_t k=2;k<m.nodes.size();++k) maxv=std::max(maxv,m.nodes[k].var);
    std::vector<bool> vars(maxv+2,false); m.support(a.id(),seen,vars);
    int r=1; for(int v=maxv+1; v>=0; --v) { if(vars[v]) r=m.mk(v,0,r); } return bdd(r); }
inline int bdd_scanset(const bdd& a, int*& varset, int& varnum) {
    minibuddy::Manager& m=minibuddy::Manager::get(); std::vector<int> vs;
    for(int n=a.id(); n>=2; n=m.nodes[n].high) vs.push_back(m.nodes[n].var);
    varnum=int(vs.size());
    if(varnum==0) { varset=0; return 0; }
    varset=static_cast<int*>(std::malloc(sizeof(int)*varnum));
    for(int i=0;i<varnum;++i) varset[i]=vs[i]; return 0; }
inline bddPair* bdd_newpair() { return new bddPair(); }
inline void bdd_freepair(bddPair* p) { delete p; }
inline int bdd_setpair(bddPair* p, int oldvar, int newvar) {
    if(int(p->map.size())<=oldvar) p->map.resize(oldvar+1,-1); p->map[oldvar]=newvar; return 0; }
inline int bdd_setpairs(bddPair* p, int* oldvars, int* newvars, int n) {
    for(int i=0;i<n;++i) bdd_setpair(p,oldvars[i],newvars[i]); return 0; }
inline bdd bdd_replace(const bdd& a, bddPair* p) {
    std::unordered_map<int,int> memo; return bdd(minibuddy::Manager::get().replace(a.id(),p->map,memo)); }
inline std::ostream& operator<<(std::ostream& os, const bdd& b) {
    minibuddy::Manager& m=minibuddy::Manager::get();
    if(b.id()==0) return os << "F"; if(b.id()==1) return os << "T";
    return os << "<" << m.nodes[b.id()].var << ":" << bdd(m.nodes[b.id()].low) << "," << bdd(m.nodes[b.id()].high) << ">"; }
#endif
