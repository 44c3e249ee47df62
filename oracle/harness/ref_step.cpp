// Oracle tooling only. Single-steps the reference calculus (TaylorCalculus, unmodified reference library) on one
// box of the van der Pol system and prints every intermediate Taylor model in hex, for bit-level debugging of
// the kernel restatement.  Usage: ref_step l0 u0 l1 u1 hmax
#include <cstdio>
#include <cstdlib>
#include "ariadne.h"
#include "taylor_calculus.h"
#include "taylor_function.h"
#include "hybrid_evolver-image.h"
using namespace Ariadne;

static void dump(const char* name, const TaylorModel& t) {
    printf("%s as=%u n=%u e=%a sweep=%g deg=%u\n", name, t.argument_size(), (unsigned)t.number_of_nonzeros(), t.error(), t.sweep_threshold(), t.maximum_degree());
    for(TaylorModel::const_iterator it=t.begin(); it!=t.end(); ++it) {
        printf("  ["); for(uint j=0;j<t.argument_size();++j) printf("%u ",(unsigned)it->key()[j]); printf("] %a\n", (double)it->data()); }
}
static void dumpv(const char* name, const Vector<TaylorModel>& v) { for(uint i=0;i<v.size();++i) { char b[64]; sprintf(b,"%s[%u]",name,i); dump(b,v[i]); } }

int main(int argc, char** argv) {
    HybridIOAutomaton a("vanderpol");
    DiscreteLocation loc("loc");
    RealParameter mu("mu",Interval(1.0,1.0));
    RealVariable x("x"), y("y");
    a.add_internal_var(x); a.add_internal_var(y);
    a.new_mode(loc);
    RealExpression dx = y;
    RealExpression dy = mu * (1.0 - Ariadne::sqr(x))*y - x;
    a.set_dynamics(loc,x,dx); a.set_dynamics(loc,y,dy);
    Box bx(2, strtod(argv[1],0), strtod(argv[2],0), strtod(argv[3],0), strtod(argv[4],0));
    double hmax=strtod(argv[5],0);
    TaylorCalculus tc;
    VectorFunction vf=a.dynamic_function(loc);
    if(argc>10) {
        // real evolver: dump the reach set models of step k
        int k=atoi(argv[10]);
        HybridBoxes domain(a.state_space(),Box(2,-3.0,3.0,-3.0,3.0));
        HybridFloatVector hmad=getHybridMidpointAbsoluteDerivatives(a,domain);
        HybridGrid grid=getHybridGrid(hmad,domain,false);
        ImageSetHybridEvolver evolver(a);
        evolver.tune_settings(domain,grid,6);
        LocalisedEnclosureType encl(loc,bx);
        std::pair<ListSet<LocalisedEnclosureType>,ListSet<LocalisedEnclosureType> > re=evolver.reach_evolve(encl,HybridTime(2.0,5),false,DIRECTION_FORWARD,UPPER_SEMANTICS);
        int i=0;
        for(ListSet<LocalisedEnclosureType>::const_iterator r=re.first.begin(); r!=re.first.end(); ++r,++i) {
            if(i==k) { dumpv("reach",r->second.models()); Box rb=r->second.bounding_box(); printf("rbox %a %a %a %a\n",rb[0].lower(),rb[0].upper(),rb[1].lower(),rb[1].upper()); }
        }
        return 0;
    }
    if(argc>6) {
        // multi-step mode: mirror the no-event path of _evolution_step (src/hybrid_evolver-image.cpp:324-547,1000-1082)
        int nsteps=atoi(argv[6]); double T=strtod(argv[7],0);
        Vector<Float> refw(2); refw[0]=strtod(argv[8],0); refw[1]=strtod(argv[9],0);
        TaylorSet set=tc.set_model(bx);
        TaylorModel time=tc.time_model(0.0,Box(2));
        for(int s=0;s<nsteps;++s) {
            printf("STEP %d\n",s);
            dumpv("set",set.models()); dump("time",time);
            Float remaining=T-time.range().lower();
            Float hp=min(remaining,hmax);
            Box range=set.range(); Float h; Box bb; Float dmax=100;
            make_lpair(h,bb)=tc.flow_bounds(vf,range,hp,dmax);
            printf("h %a bb %a %a %a %a\n",h,bb[0].lower(),bb[0].upper(),bb[1].lower(),bb[1].upper());
            TaylorCalculus::FlowModelType fm=tc.flow_model(vf,range,h,bb);
            dumpv("flow",fm.models());
            ScalarTaylorFunction idt=ScalarTaylorFunction::variable(Box(1u,Interval(-h,+h)),0u);
            TaylorSet fset=unchecked_apply(fm,combine(set.models(),idt.model()));
            dumpv("fset",fset.models());
            uint p=set.argument_size();
            TaylorModel zt=tc.time_model(0.0,Box(p)); TaylorModel ot=tc.time_model(1.0,Box(p));
            TaylorSet reach=tc.reachability_step(fset,zt,ot);
            dumpv("reach",reach.models());
            { Box rb=reach.bounding_box(); printf("rbox %a %a %a %a\n",rb[0].lower(),rb[0].upper(),rb[1].lower(),rb[1].upper()); }
            TaylorSet ev=tc.integration_step(fset,ot);
            dumpv("evolved",ev.models());
            TaylorModel ft=time+ot*h;
            ev.uniform_error_recondition(refw/4);
            int diff=ev.argument_size()-ft.argument_size();
            if(diff>0) ft=embed(ft,diff);
            Array<uint> disc=ev.kuhn_recondition();
            if(!disc.empty()) ft=recondition(ft,disc,ev.dimension(),ev.dimension());
            set=ev; time=ft;
        }
        return 0;
    }
    TaylorSet set=tc.set_model(bx);
    dumpv("set",set.models());
    Box range=set.range();
    printf("range %a %a %a %a\n",range[0].lower(),range[0].upper(),range[1].lower(),range[1].upper());
    Float h; Box bb; Float dmax=100;
    make_lpair(h,bb)=tc.flow_bounds(vf,range,hmax,dmax);
    printf("h %a bb %a %a %a %a\n",h,bb[0].lower(),bb[0].upper(),bb[1].lower(),bb[1].upper());
    {
        VectorTaylorFunction vfm(bb,vf);
        dumpv("vfm",vfm.models());
        Vector<TaylorModel> usvf(2);
        for(uint i=0;i<2;++i) usvf[i]=vfm.models()[i]*(h/rad_ivl(bb[i]));
        dumpv("usvf",usvf);
        Vector<TaylorModel> y0=TaylorModel::scalings(range);
        Vector<TaylorModel> yz=embed(unscale(y0,bb),1u);
        dumpv("yz",yz);
        { Vector<TaylorModel> y6=unchecked_flow(usvf,yz,6); Vector<TaylorModel> sc=scale(y6,bb); dumpv("scaled",sc); Vector<TaylorModel> sp=split(sc,2,true); dumpv("splitm",sp); }
        for(uint k=0;k<=6;++k) { Vector<TaylorModel> yk=unchecked_flow(usvf,yz,k); char b[32]; sprintf(b,"y%u",k); dumpv(b,yk); }
    }
    TaylorCalculus::FlowModelType fm=tc.flow_model(vf,range,h,bb);
    dumpv("flow",fm.models());
    ScalarTaylorFunction idt=ScalarTaylorFunction::variable(Box(1u,Interval(-h,+h)),0u);
    TaylorSet fset=unchecked_apply(fm,combine(set.models(),idt.model()));
    dumpv("fset",fset.models());
    TaylorModel zt=tc.time_model(0.0,Box(2)); TaylorModel ot=tc.time_model(1.0,Box(2));
    TaylorSet reach=tc.reachability_step(fset,zt,ot);
    dumpv("reach",reach.models());
    TaylorSet ev=tc.integration_step(fset,ot);
    dumpv("evolved",ev.models());
    return 0;
}
