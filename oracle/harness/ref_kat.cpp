// Oracle tooling only (test infrastructure; never linked into the product). Known-answer vectors for the PRIMITIVES of
// the hot path, produced by the unmodified reference library: TaylorModel operators / functions
// (include/taylor_model.h, src/taylor_model.cpp) and Interval arithmetic (include/numeric.h, src/numeric.cpp).
// Inputs are (a) the literal vectors of the reference's own unit tests -- test/test_taylor_model.cpp:164-177 (arithmetic),
// :224-234 (intersection), :236-249 (split), :252-270 (antiderivative), :197-199 (rec / sqrt), test/test_interval.cpp:152-254 --
// and (b) deterministic pseudo-random sparse models in 2 and 3 variables whose coefficients do round, so that the
// upward-rounded error chains of _add2 / _mul4 / _scal / antidifferentiate are pinned bit for bit.
// Output (hex floats) is read by tests/test_gpu_kat.py, which replays every line on the device (ab200_taylor_model_op,
// ab200_interval_op) and compares bit-exactly.
//   ref_kat > tests/golden/kat_primitives.txt
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ariadne.h"
#include "taylor_model.h"
using namespace Ariadne;
typedef TaylorModel T;
typedef Expansion<Float> E;

static void dump_tm(const char* tag, const T& t) {
    printf("%s %u %a\n", tag, (unsigned)t.number_of_nonzeros(), (double)t.error());
    for(T::const_iterator it=t.begin(); it!=t.end(); ++it) {
        printf("t"); for(uint j=0;j<t.argument_size();++j) printf(" %u",(unsigned)it->key()[j]); printf(" %a\n",(double)it->data()); }
}
static void head(const char* op, uint as, int k, int flag, double s, Interval iv) {
    printf("tm %s as %u k %d flag %d s %a il %a iu %a\n", op, as, k, flag, s, iv.lower(), iv.upper());
}
static T none(uint as) { return T(as); }

// one line of every binary / unary operation the device implements, on the given operands
static void emit(const char* op, const T& x, const T& y, int k, int flag, double s, Interval iv) {
    std::string o(op);
    set_rounding_to_nearest();
    T r(x.argument_size());
    Interval rng(0,0); bool is_range=false;
    try {
        if(o=="add") r=x+y; else if(o=="sub") r=x-y; else if(o=="mul") r=x*y;
        else if(o=="addf") r=x+Float(s); else if(o=="subf") r=x-Float(s); else if(o=="mulf") r=x*Float(s); else if(o=="divf") r=x/Float(s);
        else if(o=="addi") r=x+iv; else if(o=="subi") r=x-iv; else if(o=="muli") r=x*iv; else if(o=="divi") r=x/iv;
        else if(o=="neg") r=-x;
        else if(o=="intersection") r=intersection(x,y);
        else if(o=="split") r=split(x,(uint)k,tribool(flag!=0));
        else if(o=="antiderivative") r=antiderivative(x,(uint)k);
        else if(o=="rec") r=rec(x); else if(o=="sqrt") r=sqrt(x); else if(o=="abs") r=abs(x); else if(o=="max") r=max(x,y);
        else if(o=="exp") r=exp(x); else if(o=="log") r=log(x); else if(o=="sin") r=sin(x); else if(o=="cos") r=cos(x);
        else if(o=="range") { rng=x.range(); is_range=true; }
        else { fprintf(stderr,"unknown op %s\n",op); exit(2); }
    } catch(std::exception& e) { set_rounding_to_nearest(); fprintf(stderr,"skipping %s: %s\n",op,e.what()); return; }
    set_rounding_to_nearest();
    head(op,x.argument_size(),k,flag,s,iv);
    dump_tm("x",x); dump_tm("y",y);
    if(is_range) printf("v %a %a\n",rng.lower(),rng.upper()); else dump_tm("r",r);
}

static unsigned long long lcg_state=0x9E3779B97F4A7C15ull;
static double rnd() { lcg_state=lcg_state*6364136223846793005ull+1442695040888963407ull; return ((double)((lcg_state>>11)&((1ull<<53)-1)))/(double)(1ull<<53)*2.0-1.0; }
// sparse pseudo-random model: every monomial of degree <= deg is present with probability `fill`; coefficients decay with the degree
static T random_model(uint as, uint deg, double fill, double scale, double err) {
    E e(as);
    for(MultiIndex a(as); a.degree()<=deg; ++a) {
        double keep=rnd(); double v=rnd()*scale; for(uint d=0; d<a.degree(); ++d) v*=0.37;
        if((keep+1.0)/2.0<fill && v!=0.0) e.append(a,v);
    }
    return T(e,err);
}

static void iv(const char* op, Interval a, Interval b, int m) {
    set_rounding_to_nearest();
    std::string o(op); Interval r(0,0);
    try {
    if(o=="add") r=a+b; else if(o=="sub") r=a-b; else if(o=="mul") r=a*b; else if(o=="div") r=a/b;
    else if(o=="addf") r=a+Float(b.lower()); else if(o=="subf") r=a-Float(b.lower()); else if(o=="mulf") r=a*Float(b.lower()); else if(o=="divf") r=a/Float(b.lower());
    else if(o=="fsub") r=Float(a.lower())-b; else if(o=="fdiv") r=Float(a.lower())/b;
    else if(o=="sqr") r=sqr(a); else if(o=="sqrt") r=sqrt(a); else if(o=="pow") r=pow(a,(uint)m); else if(o=="rec") r=rec(a);
    else if(o=="abs") r=abs(a); else if(o=="neg") r=-a; else if(o=="max") r=max(a,b); else if(o=="min") r=min(a,b);
    else if(o=="exp") r=exp(a); else if(o=="log") r=log(a); else if(o=="sin") r=sin(a); else if(o=="cos") r=cos(a);
    else { fprintf(stderr,"unknown interval op %s\n",op); exit(2); }
    } catch(std::exception& e) { set_rounding_to_nearest(); fprintf(stderr,"skipping interval %s: %s\n",op,e.what()); return; }   // e.g. the reference's log of a point below 1 gives lower>upper
    set_rounding_to_nearest();
    printf("iv %s a %a %a b %a %a m %d r %a %a\n",op,a.lower(),a.upper(),b.lower(),b.upper(),m,r.lower(),r.upper());
}

int main() {
    const Interval I0(0,0);
    // ---- test/test_taylor_model.cpp:164-177 ----
    T a(E(1,2, 1.0,-2.0,3.0),0.75), b(E(1,2, 3.0,2.0,-4.0),0.5), c(E(1,2, 0.0,0.0,3.0),0.75);
    emit("addf",a,none(1),0,0,-3.0,I0); emit("subf",a,none(1),0,0,-3.0,I0); emit("mulf",a,none(1),0,0,-3.0,I0); emit("divf",a,none(1),0,0,-4.0,I0);
    emit("addi",a,none(1),0,0,0.0,Interval(-1,2)); emit("subi",a,none(1),0,0,0.0,Interval(-1,2)); emit("muli",a,none(1),0,0,0.0,Interval(-1,2));
    emit("divi",a,none(1),0,0,0.0,Interval(0.25,2.0)); emit("neg",a,none(1),0,0,0.0,I0);
    emit("add",a,b,0,0,0.0,I0); emit("sub",a,b,0,0,0.0,I0); emit("mul",c,b,0,0,0.0,I0); emit("mul",a,b,0,0,0.0,I0);
    // ---- :224-234 intersection ----
    emit("intersection",T(E(1,4, 1.0,-0.75,0.0,3.0,3.25),0.5),T(E(1,4, 1.0,0.0,0.25,2.0,3.0),1.0),0,0,0.0,I0);
    emit("intersection",T(E(1,0, 2./3),0.5),T(E(1,0, 6./5),0.25),0,0,0.0,I0);
    // ---- :236-249 split ----
    { T x=T::variable(2,0), y=T::variable(2,1); T t=1+3*x+2*y-5*x*x-7*x*y+11*y*y;
      emit("split",t,none(2),0,0,0.0,I0); emit("split",t,none(2),0,1,0.0,I0); emit("split",t,none(2),1,0,0.0,I0); emit("split",t,none(2),1,1,0.0,I0); }
    // ---- :252-270 antiderivative ----
    emit("antiderivative",T(E(2,1, 0,0,2.0),0.),none(2),0,0,0.0,I0); emit("antiderivative",T(E(2,1, 0,0,2.0),0.),none(2),1,0,0.0,I0);
    emit("antiderivative",T(E(2,1, 1,0,3.0),0.),none(2),0,0,0.0,I0); emit("antiderivative",T(E(2,1, 2,0,7.5),0.),none(2),0,0,0.0,I0);
    emit("antiderivative",T(E(2,1, 2,4,7.5),0.),none(2),0,0,0.0,I0); emit("antiderivative",T(E(2,1, 2,4,7.5),0.),none(2),1,0,0.0,I0);
    emit("antiderivative",T(E(1,1, 2,2.0),0.),none(1),0,0,0.0,I0); emit("antiderivative",T(E(1,1, 2,2.0),1.),none(1),0,0,0.0,I0);
    emit("antiderivative",T(E(2,6, 0,0,1., 1,0,2., 0,1,3., 2,0,4., 1,1,5., 0,2,6.),0.),none(2),0,0,0.0,I0);
    // ---- :197-199 functions the hot path uses (rec, sqrt) on xo = 1 + x/2 ----
    { T xo(E(1,1, 1.0,0.5),0.0); emit("rec",xo,none(1),0,0,0.0,I0); emit("sqrt",xo,none(1),0,0,0.0,I0); emit("range",xo,none(1),0,0,0.0,I0); }
    // ---- pseudo-random sparse models: coefficients that round ----
    for(uint as=2; as<=3; ++as) for(int rep=0; rep<3; ++rep) {
        const uint deg=(as==2)?6:4;
        T x=random_model(as,deg,0.8,1.0,1e-9*(rep+1)), y=random_model(as,deg,0.6,0.7,3e-10);
        emit("add",x,y,0,0,0.0,I0); emit("sub",x,y,0,0,0.0,I0); emit("mul",x,y,0,0,0.0,I0); emit("mul",y,x,0,0,0.0,I0); emit("mul",x,x,0,0,0.0,I0);
        emit("mulf",x,none(as),0,0,rnd()*3.0,I0); emit("divf",x,none(as),0,0,0.3+rnd()*0.1,I0); emit("addf",x,none(as),0,0,rnd(),I0);
        { double l=rnd(); emit("muli",x,none(as),0,0,0.0,Interval(l,l+0.25)); emit("addi",x,none(as),0,0,0.0,Interval(l,l+1e-3)); emit("divi",x,none(as),0,0,0.0,Interval(1.5+l,2.75+l)); }
        emit("neg",x,none(as),0,0,0.0,I0); emit("range",x,none(as),0,0,0.0,I0);
        emit("intersection",x+Interval(-1e-3,1e-3),x+y*Float(1e-5)+Interval(-2e-3,2e-3),0,0,0.0,I0);
        for(uint k=0;k<as;++k) { emit("split",x,none(as),(int)k,0,0.0,I0); emit("split",x,none(as),(int)k,1,0.0,I0); emit("antiderivative",x,none(as),(int)k,0,0.0,I0); }
        T p=random_model(as,3,0.9,0.2,1e-8)+Float(2.0);
        emit("rec",p,none(as),0,0,0.0,I0); emit("sqrt",p,none(as),0,0,0.0,I0);
        T q=random_model(as,2,0.9,0.5,1e-8);
        emit("abs",q,none(as),0,0,0.0,I0); emit("max",x,y,0,0,0.0,I0); emit("abs",p,none(as),0,0,0.0,I0);
    }
    // ---- exp / log / sin / cos of models: the vectors of test/test_taylor_model.cpp:183-205 and rounding pseudo-random models ----
    { T x(E(1,1, 0.0,1.0),0.0), xz(E(1,1, 0.0,0.5),0.0), xo(E(1,1, 1.0,0.5),0.0), xl(E(1,1, 0.693147,0.5),0.0);
      emit("exp",x,none(1),0,0,0.0,I0); emit("sin",x,none(1),0,0,0.0,I0); emit("cos",x,none(1),0,0,0.0,I0);
      emit("exp",xz,none(1),0,0,0.0,I0); emit("sin",xz,none(1),0,0,0.0,I0); emit("cos",xz,none(1),0,0,0.0,I0);
      emit("log",xo,none(1),0,0,0.0,I0); emit("exp",xl,none(1),0,0,0.0,I0); }
    for(uint as=2; as<=3; ++as) for(int rep=0; rep<2; ++rep) {
        T q=random_model(as,3,0.9,0.4,1e-8), p=random_model(as,3,0.9,0.2,1e-8)+Float(2.0), w=random_model(as,2,0.9,1.5,1e-6)+Float(7.0+3*rep);
        emit("exp",q,none(as),0,0,0.0,I0); emit("sin",q,none(as),0,0,0.0,I0); emit("cos",q,none(as),0,0,0.0,I0);
        emit("exp",w,none(as),0,0,0.0,I0); emit("sin",w,none(as),0,0,0.0,I0); emit("cos",w,none(as),0,0,0.0,I0);
        emit("log",p,none(as),0,0,0.0,I0); emit("log",w,none(as),0,0,0.0,I0);
    }
    // ---- Interval arithmetic: test/test_interval.cpp:152-254 ----
    const double eps=2.2204460492503131e-16, mn=4.9406564584124654e-324;
    iv("add",Interval(1.5),Interval(mn),0); iv("sub",Interval(1.5),Interval(mn),0);
    iv("mul",Interval(1+eps,1+2*eps),Interval(1+eps,1+3*eps),0); iv("div",Interval(1),Interval(3),0); iv("div",Interval(2),Interval(5),0);
    iv("addf",Interval(1.5),Interval(mn),0); iv("subf",Interval(1.5),Interval(mn),0); iv("mulf",Interval(1+eps,1+2*eps),Interval(1+eps),0);
    iv("divf",Interval(1+3*eps,1+5*eps),Interval(1+eps),0); iv("fsub",Interval(mn),Interval(1.5),0); iv("fdiv",Interval(1+5*eps),Interval(1+2*eps,1+3*eps),0);
    iv("sqr",Interval(1-eps,1+eps),I0,0); iv("rec",Interval(1+2*eps,1+5*eps),I0,0);
    iv("add",Interval(5,7),Interval(2,4),0); iv("sub",Interval(5,7),Interval(2,6),0);
    { const double s[3][2]={{5,7},{-7,5},{-7,-5}}, t[3][2]={{2,4},{-2,4},{-4,-2}};
      for(int i=0;i<3;++i) for(int j=0;j<3;++j) iv("mul",Interval(s[i][0],s[i][1]),Interval(t[j][0],t[j][1]),0);
      for(int i=0;i<3;++i) for(int j=0;j<3;j+=2) iv("div",Interval(s[i][0],s[i][1]),Interval(t[j][0],t[j][1]),0);
      const double p[4][2]={{5,7},{-5,7},{-7,5},{-7,-5}}; const int ms[7]={0,1,2,3,4,5,7};
      for(int q=0;q<7;++q) for(int i=0;i<4;++i) iv("pow",Interval(p[i][0],p[i][1]),I0,ms[q]); }
    // values that round: thirds, fifths, sevenths
    for(int i=1;i<=6;++i) { Interval x(1.0/(2*i+1),2.0/(2*i+1)), y(-3.0/7*i,5.0/(3*i));
        iv("add",x,y,0); iv("sub",x,y,0); iv("mul",x,y,0); iv("mul",y,y,0); iv("div",y,x,0); iv("sqr",y,I0,0); iv("sqr",x,I0,0); iv("sqrt",x,I0,0); iv("rec",x,I0,0);
        iv("pow",y,I0,3); iv("pow",y,I0,4); iv("abs",y,I0,0); iv("neg",y,I0,0); iv("max",x,y,0); iv("min",x,y,0); iv("mulf",y,Interval(1.0/3),0); iv("divf",y,Interval(7.0/3),0); }
    // ---- exp / log / sin / cos of intervals (src/numeric.cpp:912-966; the reference's own rounded series, no libm) ----
    { const double xs[]={ -20.5, -3.25, -1.0, -0.4, -0.1, -1e-3, 0.0, 1e-9, 0.05, 0.3, 0.5, 0.7, 1.0, 1.5, 2.0, 2.5, 3.0, 3.14, 3.2, 4.0, 4.7, 5.5, 6.2, 6.3, 7.0, 9.5, 12.0, 25.0, 100.25, 710.0 };
      const int nx=sizeof(xs)/sizeof(xs[0]);
      for(int i=0;i<nx;++i) { double x=xs[i]; double w=(i%3==0)?0.0:((i%3==1)?1e-3:0.37);
          iv("exp",Interval(x<700?x:700.0,(x<700?x:700.0)+w),I0,0);
          iv("sin",Interval(x,x+w),I0,0); iv("cos",Interval(x,x+w),I0,0); iv("sin",Interval(-x-w,-x),I0,0); iv("cos",Interval(-x-w,-x),I0,0);
          if(x>0) { iv("log",Interval(x,x+w),I0,0); iv("log",Interval(x/1024,x/1024+w/7),I0,0); } }
      iv("cos",Interval(-1.0,2.0),I0,0); iv("cos",Interval(1.0,4.0),I0,0); iv("cos",Interval(3.0,7.0),I0,0); iv("cos",Interval(-0.5,13.0),I0,0);
      iv("sin",Interval(-1.0,2.0),I0,0); iv("sin",Interval(1.0,4.0),I0,0); iv("exp",Interval(-1.0/3,1.0/3),I0,0); iv("log",Interval(1.0/3,3.0),I0,0); }
    return 0;
}
