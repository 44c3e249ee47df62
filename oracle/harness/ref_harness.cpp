// Oracle tooling only (never linked into the product). Drives the UNMODIFIED reference library
// (oracle/_ref/libariadne_ref.a, built by oracle/Makefile from /root/reference/src) through its public API:
//   HybridReachabilityAnalyser::upper_reach_evolve   (reference: src/reachability_analyser.cpp:309-377)
//   HybridReachabilityAnalyser::outer_chain_reach    (reference: src/reachability_analyser.cpp:380-493)
//   EvolverInterface::reach_evolve                   (reference: include/evolver_base.h:77-78)
// and writes golden vectors (hex floats, bit-exact) or timings.
//
// Usage:  ref_harness <config> <mode> [key=value ...]
//   config: watertank | bouncingball | vanderpol | springs | boost | spiral
//   mode:   info      print grid / lock time / evolver tuning constants
//           cells     per-enclosure reach/final bounding boxes for the first lap's frontier (parity goldens)
//           reach     full analyser call; dumps reach (and evolve) cell boxes per location
//           time      full analyser call, wall clock + cells evolved (CPU baseline)
//   keys:   acc=<int> T=<float> events=<int> ncells=<int> out=<file> verb=<int> laps=<int>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <map>
#include <chrono>
#include <sstream>
#include <fstream>

#include "ariadne.h"
#include "taylor_calculus.h"
#include "hybrid_evolver-image.h"
#include "logging.h"
#include "integrator.h"
#include "taylor_function.h"
#include <boost/thread.hpp>
#include "workers.h"
#include "watertank-hysteresis.h"
#include "springs.h"
#include "boost-converter.h"
#include "workers_b200.h"
#include "evolver_b200.h"

using namespace Ariadne;

typedef LocalisedEnclosureType HEncl;
typedef ListSet<LocalisedEnclosureType> HEnclList;

struct Config {
    std::string name;
    HybridIOAutomaton sys;
    HybridBoxes domain;
    HybridBoundedConstraintSet initial;
    int acc;
    double T;
    int events;
    bool chain;
    std::string cons_text;      // the constraint-set initial set as text for ab200_analyser_discretise_constraint (cons=...)
};

// van der Pol oscillator, mu=1: x'=y, y'=mu(1-x^2)y-x  (system of examples/vanderpol.cc:36-54)
static HybridIOAutomaton make_vanderpol() {
    HybridIOAutomaton a("vanderpol");
    DiscreteLocation loc("loc");
    RealParameter mu("mu",Interval(1.0,1.0));
    RealVariable x("x"), y("y");
    a.add_internal_var(x); a.add_internal_var(y);
    a.new_mode(loc);
    RealExpression dx = y;
    RealExpression dy = mu * (1.0 - Ariadne::sqr(x))*y - x;
    a.set_dynamics(loc,x,dx);
    a.set_dynamics(loc,y,dy);
    return a;
}

// damped pendulum with transcendental terms (not a BASELINE config: it exists to pin exp / log / sin / cos on the tapes):
// x'=y, y'=-sin(x)+0.1 cos(y)-0.05 exp(0.3 x)+0.01 log(12+x^2)
static HybridIOAutomaton make_pendulum() {
    HybridIOAutomaton a("pendulum");
    DiscreteLocation loc("loc");
    RealVariable x("x"), y("y");
    a.add_internal_var(x); a.add_internal_var(y);
    a.new_mode(loc);
    RealExpression dx = y;
    RealExpression dy = -Ariadne::sin(x) + 0.1*Ariadne::cos(y) - 0.05*Ariadne::exp(0.3*x) + 0.01*Ariadne::log(12.0+Ariadne::sqr(x));
    a.set_dynamics(loc,x,dx);
    a.set_dynamics(loc,y,dy);
    return a;
}

// bouncing ball: x'=vx, vx'=-g; forced x<=0 -> vx:=-a vx  (system of examples/bouncingball.cc:33-64)
static HybridIOAutomaton make_bouncingball() {
    RealParameter a("a",0.5), g("g",9.81);
    HybridIOAutomaton ball("ball");
    DiscreteLocation freefall("freefall");
    DiscreteEvent bounce("bounce");
    RealVariable x("x"), vx("vx");
    ball.add_internal_var(x); ball.add_internal_var(vx);
    ball.new_mode(freefall);
    RealExpression dx = vx;
    RealExpression dvx = -g;
    ball.set_dynamics(freefall,x,dx);
    ball.set_dynamics(freefall,vx,dvx);
    RealExpression guard = -x;
    std::map<RealVariable,RealExpression> reset;
    reset[x] = x;
    reset[vx] = -a * vx;
    ball.new_forced_transition(bounce,freefall,freefall,reset,guard);
    return ball;
}

// the vector field of test/test_integrator.cpp:73-157: x'=2, y'=3
static HybridIOAutomaton make_constder() {
    HybridIOAutomaton a("constder");
    DiscreteLocation loc("loc");
    RealVariable x("x"), y("y");
    a.add_internal_var(x); a.add_internal_var(y);
    a.new_mode(loc);
    RealExpression dx = 2.0;
    RealExpression dy = 3.0;
    a.set_dynamics(loc,x,dx);
    a.set_dynamics(loc,y,dy);
    return a;
}

static std::map<std::string,std::string> parse_kv(int argc, char** argv, int first) {
    std::map<std::string,std::string> kv;
    for(int i=first;i<argc;++i) { std::string s(argv[i]); size_t p=s.find('='); if(p!=std::string::npos) kv[s.substr(0,p)]=s.substr(p+1); }
    return kv;
}
static double getd(std::map<std::string,std::string>& kv, const char* k, double d) { return kv.count(k)?atof(kv[k].c_str()):d; }
static int geti(std::map<std::string,std::string>& kv, const char* k, int d) { return kv.count(k)?atoi(kv[k].c_str()):d; }

static Box box_from(std::map<std::string,std::string>& kv, const char* key, const Box& dflt) {
    if(!kv.count(key)) return dflt;
    std::vector<double> v; std::stringstream ss(kv[key]); std::string tok;
    while(std::getline(ss,tok,',')) v.push_back(atof(tok.c_str()));
    Box b(v.size()/2); for(uint i=0;i<v.size()/2;++i) b[i]=Interval(v[2*i],v[2*i+1]); return b;
}

static Config make_config(const std::string& name, std::map<std::string,std::string>& kv) {
    Config c; c.name=name; c.events=geti(kv,"events",5); c.chain=false;
    if(name=="vanderpol") {
        c.sys=make_vanderpol();
        c.domain=HybridBoxes(c.sys.state_space(),box_from(kv,"domain",Box(2,-3.0,3.0,-3.0,3.0)));
        c.initial=HybridBoundedConstraintSet(c.sys.state_space());
        c.initial[DiscreteLocation("loc")]=box_from(kv,"init",Box(2,-2.2,2.2,-2.8,2.8));
        c.acc=geti(kv,"acc",6); c.T=getd(kv,"T",0.0);
    } else if(name=="pendulum") {
        c.sys=make_pendulum();
        c.domain=HybridBoxes(c.sys.state_space(),box_from(kv,"domain",Box(2,-3.0,3.0,-3.0,3.0)));
        c.initial=HybridBoundedConstraintSet(c.sys.state_space());
        c.initial[DiscreteLocation("loc")]=box_from(kv,"init",Box(2,0.9,1.2,-0.2,0.1));
        c.acc=geti(kv,"acc",4); c.T=getd(kv,"T",0.0);
    } else if(name=="constder") {
        c.sys=make_constder();
        c.domain=HybridBoxes(c.sys.state_space(),box_from(kv,"domain",Box(2,-1.0,5.0,-1.0,6.0)));
        c.initial=HybridBoundedConstraintSet(c.sys.state_space());
        c.initial[DiscreteLocation("loc")]=box_from(kv,"init",Box(2,0.0,1.0,-0.5,1.5));
        c.acc=geti(kv,"acc",3); c.T=getd(kv,"T",0.0);
    } else if(name=="bouncingball") {
        c.sys=make_bouncingball();
        // state order is alphabetical: (vx,x)
        c.domain=HybridBoxes(c.sys.state_space(),box_from(kv,"domain",Box(2,-10.1,10.1,-0.1,2.1)));
        c.initial=HybridBoundedConstraintSet(c.sys.state_space());
        c.initial[DiscreteLocation("freefall")]=box_from(kv,"init",Box(2,0.0,0.001,1.999,2.0));
        c.acc=geti(kv,"acc",6); c.T=getd(kv,"T",0.0); c.chain=true;
    } else if(name=="watertank") {
        c.sys=getWatertankHysteresis();
        c.domain=HybridBoxes(c.sys.state_space(),box_from(kv,"domain",Box(2,0.0,1.0,4.5,9.0)));
        c.initial=HybridBoundedConstraintSet(c.sys.state_space());
        c.initial[DiscreteLocation("flow,idle,rising")]=box_from(kv,"initr",Box(2,1.0,1.0,6.0,7.5));
        c.initial[DiscreteLocation("flow,idle,falling")]=box_from(kv,"initf",Box(2,0.0,0.0,6.0,7.5));
        c.acc=geti(kv,"acc",6); c.T=getd(kv,"T",0.0);
    } else if(name=="springs") {
        c.sys=getSpringsSystem();
        c.domain=HybridBoxes(c.sys.state_space(),box_from(kv,"domain",Box(5,0.0,40.0,-2.0,2.0,-3.0,3.0,-1.0,3.0,0.0,4.0)));
        c.initial=HybridBoundedConstraintSet(c.sys.state_space());
        c.initial[DiscreteLocation("free1,free2")]=box_from(kv,"init",Box(5,0.0,0.0,0.0,0.0,0.0,0.0,0.0,0.0,3.0,3.0));
        c.acc=geti(kv,"acc",3); c.T=getd(kv,"T",0.0);
    } else if(name=="boost") {
        c.sys=getBoostConverter();
        c.domain=HybridBoxes(c.sys.state_space(),box_from(kv,"domain",Box(3,0.0,0.001,0.0,3.0,0.0,6.0)));
        c.initial=HybridBoundedConstraintSet(c.sys.state_space());
        c.initial[DiscreteLocation("incr,below_duty")]=box_from(kv,"init",Box(3,0.0,0.0,1.0,1.0,1.0,1.0));
        c.acc=geti(kv,"acc",4); c.T=getd(kv,"T",0.0); c.chain=true;
    } else {
        fprintf(stderr,"unknown config %s\n",name.c_str()); exit(2);
    }
    if(kv.count("chain")) c.chain=geti(kv,"chain",0)!=0;
    if(kv.count("cons")) {
        // Constraint-set initial set in the first location: { x in init box : f(x) in codomain } with
        //   cons=disc   f = (x0-cx)^2 + (x1-cy)^2, codomain [-1, r^2]       (centre = cc=cx,cy; r = cr)
        //   cons=band   f = (x0 + x1, x0*x1),     codomain [slo,shi] x [plo,phi]  (cb=slo,shi,plo,phi)
        DiscreteLocation loc=c.sys.modes().front().location();
        RealSpace sp=c.sys.continuous_state_space(loc);
        Box dom=c.initial[loc].domain();
        List<RealExpression> fs; Box cod;
        RealExpression x0(sp.variable(0)), x1(sp.variable(1));
        if(kv["cons"]=="disc") {
            Box cc=box_from(kv,"cc",Box(1,1.2,1.3)); double r=getd(kv,"cr",0.25);
            fs.append(Ariadne::sqr(x0-cc[0].lower())+Ariadne::sqr(x1-cc[0].upper()));
            cod=Box(1,-1.0,r*r);
        } else {
            Box cb=box_from(kv,"cb",Box(2,2.2,2.6,1.3,1.7));
            fs.append(x0+x1); fs.append(x0*x1);
            cod=cb;
        }
        c.initial[loc]=BoundedConstraintSet(dom,VectorFunction(fs,sp),cod);
        c.cons_text="constraints\n"+b200_constraint_text(c.sys,loc,dom,fs,cod);
    }
    return c;
}

// ---- system -> op-tape text: adapter/workers_b200.cpp (b200_system_text), the same walker the adapter ships ----
static void dump_system(FILE* f, const HybridIOAutomaton& sys) { std::string t=b200_system_text(sys); fputs(t.c_str(),f); }

static void dump_box(FILE* f, const Box& b) { for(uint i=0;i<b.size();++i) fprintf(f," %a %a",b[i].lower(),b[i].upper()); }

static void dump_set(FILE* f, const char* tag, const HybridDenotableSet& s) {
    for(HybridDenotableSet::locations_const_iterator li=s.locations_begin(); li!=s.locations_end(); ++li) {
        fprintf(f,"%s loc %s count %lu height %u\n",tag,li->first.name().c_str(),(unsigned long)li->second.size(),li->second.root_cell_height());
        for(DenotableSetType::const_iterator ci=li->second.begin(); ci!=li->second.end(); ++ci) {
            fprintf(f,"c"); dump_box(f,*ci); fprintf(f,"\n");
        }
    }
}

int main(int argc, char** argv) {
    if(argc<3) { fprintf(stderr,"usage: ref_harness <config> <mode> [key=value...]\n"); return 2; }
    std::map<std::string,std::string> kv=parse_kv(argc,argv,3);
    Config c=make_config(argv[1],kv);
    if(kv.count("chain")) c.chain=geti(kv,"chain",0)!=0;
    std::string mode(argv[2]);
    // midparams=1: every nonsingleton parameter replaced by its midpoint, as epsilon_lower_chain_reach substitutes them
    // (getMidpointsSet, src/reachability_analyser.cpp:1000-1014) -- used with mode "system" for the substituted system's text
    if(geti(kv,"midparams",0)) { std::list<RealParameterSet> l; l.push_back(c.sys.nonsingleton_parameters()); c.sys.substitute_all(getMidpointsSet(l).front()); }
    FILE* out=stdout; if(kv.count("out")) out=fopen(kv["out"].c_str(),"w");
    int verb=geti(kv,"verb",0);

    if(mode=="system") { dump_system(out,c.sys); return 0; }
    if(mode=="integrator") {
        // IntegratorInterface::flow_bounds / flow of the reference's TaylorIntegrator (src/integrator.cpp:44-152) on the dynamics
        // of one location: the golden vectors of the B4 facade (ab200_integrator_flow_bounds, ab200_integrator_flow)
        DiscreteLocation loc(kv.count("loc")?kv["loc"]:c.sys.modes().front().location().name());
        VectorFunction vf=c.sys.dynamic_function(loc);
        Box bx=box_from(kv,"box",Box(2,0.0,1.0,-0.5,1.5));
        double hmax=getd(kv,"h",1.0); int order=geti(kv,"order",4);
        TaylorIntegrator integrator(order);
        Pair<Float,Vector<Interval> > fb=integrator.flow_bounds(vf,bx,hmax);
        fprintf(out,"integrator loc %s order %d hmax %a box",loc.name().c_str(),order,hmax); dump_box(out,bx); fprintf(out,"\n");
        fprintf(out,"bounds h %a b",fb.first); dump_box(out,Box(fb.second)); fprintf(out,"\n");
        VectorTaylorFunction phi=integrator.flow(vf,bx,hmax);
        fprintf(out,"domain"); dump_box(out,Box(phi.domain())); fprintf(out,"\n");
        for(uint i=0;i<phi.result_size();++i) {
            const TaylorModel& m=phi.models()[i];
            fprintf(out,"model %u terms %u error %a\n",i,(unsigned)m.number_of_nonzeros(),m.error());
            for(TaylorModel::const_iterator it=m.begin(); it!=m.end(); ++it) {
                fprintf(out,"t"); for(uint j=0;j<m.argument_size();++j) fprintf(out," %u",(unsigned)it->key()[j]); fprintf(out," %a\n",it->data());
            }
        }
        return 0;
    }

    HybridReachabilityAnalyser analyser(c.sys,c.domain,c.acc);
    analyser.verbosity=verb;
    if(kv.count("lock")) analyser.settings().lock_to_grid_time=getd(kv,"lock",0.0);
    double lock=analyser.settings().lock_to_grid_time;
    double T = c.T>0 ? c.T : getd(kv,"laps",1.0)*lock;

    HybridFloatVector hmad=getHybridMidpointAbsoluteDerivatives(c.sys,c.domain);
    HybridGrid grid=getHybridGrid(hmad,c.domain,false);
    // b200=1: the same calls through the EvolverInterface seam of the adapter (adapter/evolver_b200.h), i.e. an
    // ImageSetHybridEvolver whose _evolution runs on libariadne_b200.so
    B200ImageSetHybridEvolver evolver_b200(c.sys);
    ImageSetHybridEvolver evolver_stock(c.sys);
    ImageSetHybridEvolver& evolver = geti(kv,"b200",0) ? static_cast<ImageSetHybridEvolver&>(evolver_b200) : evolver_stock;
    // The analyser tunes its evolver on ReachabilityRestriction::bounding_box(), i.e. the domain widened by one cell
    // while the domain has not been discretised (src/reachability_analyser.cpp:125-142,
    // src/reachability_restriction.cpp:597-621); outer=0 tunes on the plain domain instead.
    if(geti(kv,"outer",1)) { ReachabilityRestriction rr(c.domain,grid,c.acc); evolver.tune_settings(rr.bounding_box(),grid,c.acc); }
    else evolver.tune_settings(c.domain,grid,c.acc);

    fprintf(out,"config %s acc %d T %a lock %a events %d chain %d\n",c.name.c_str(),c.acc,T,lock,c.events,(int)c.chain);
    for(HybridGrid::locations_const_iterator gi=grid.locations_begin(); gi!=grid.locations_end(); ++gi) {
        const Grid& g=gi->second;
        fprintf(out,"grid loc %s n %u origin",gi->first.name().c_str(),g.dimension());
        for(uint i=0;i<g.dimension();++i) fprintf(out," %a",g.origin()[i]);
        fprintf(out," lengths"); for(uint i=0;i<g.dimension();++i) fprintf(out," %a",g.lengths()[i]);
        fprintf(out," hmad"); for(uint i=0;i<g.dimension();++i) fprintf(out," %a",hmad.find(gi->first)->second[i]);
        fprintf(out," maxstep %a",evolver.settings().maximum_step_size().find(gi->first)->second);
        fprintf(out," refw"); for(uint i=0;i<g.dimension();++i) fprintf(out," %a",evolver.settings().reference_enclosure_widths().find(gi->first)->second[i]);
        fprintf(out,"\n");
    }
    if(mode=="info") return 0;

    if(mode=="cells") {
        // First-lap frontier: enumerate the discretised initial set as the analyser does, then evolve each
        // enclosure for one lock_to_grid_time through reach_evolve and dump the bounding boxes that the
        // discretiser consumes (include/hybrid_set.h:976-996).
        ReachabilityRestriction restriction(c.domain,grid,c.acc);
        HybridDenotableSet initial=restriction.outer_intersection_with(c.initial);
        int mince=geti(kv,"mince",0);
        if(mince) { for(HybridDenotableSet::locations_iterator li=initial.locations_begin(); li!=initial.locations_end(); ++li) li->second.mince(c.acc); }
        std::list<HEncl> encl=to_enclosures(initial);
        int ncells=geti(kv,"ncells",1<<30); int stride=geti(kv,"stride",1); int offset=geti(kv,"offset",0);
        bool ignore_act=geti(kv,"ignore_activations",c.chain?1:0)!=0;
        HybridTime ht(T,c.chain?1:c.events);
        evolver.verbosity=geti(kv,"everb",0);
        fprintf(out,"frontier %lu ignore_activations %d lower %d backward %d\n",(unsigned long)encl.size(),(int)ignore_act,geti(kv,"lower",0),geti(kv,"backward",0));
        HybridDenotableSet ureach(grid), ufinal(grid);
        int k=0, done=0;
        // to_enclosures (src/reachability_analyser.cpp:1244-1253) walks the set in this same order; the cell box itself
        // is printed (the enclosure's bounding box is the range of scalings(box), which is a few ulps wider).
        HybridDenotableSet::const_iterator cell_it=initial.begin();
        for(std::list<HEncl>::const_iterator it=encl.begin(); it!=encl.end() && done<ncells; ++it,++k,++cell_it) {
            if(k<offset || (k-offset)%stride) continue;
            ++done;
            // midpoints=1: the enclosure is the centre of (cell box ^ initial box), as lower_reach_evolve and
            // epsilon_lower_chain_reach start (_enclosures_from_discretised_initial_set_midpoints, src/reachability_analyser.cpp:1220-1241)
            HEncl start=*it; Box start_box=cell_it->second;
            if(geti(kv,"midpoints",0)) { Box isect=intersection(Box(cell_it->second),c.initial.find(it->first)->second.domain()); start_box=Box(isect.centre()); start=HEncl(it->first,HEncl::second_type(start_box)); }
            fprintf(out,"cell %d loc %s box",k,it->first.name().c_str()); dump_box(out,start_box); fprintf(out,"\n");
            try {
                std::pair<HEnclList,HEnclList> re=evolver.reach_evolve(start,ht,ignore_act,geti(kv,"backward",0)?DIRECTION_BACKWARD:DIRECTION_FORWARD,geti(kv,"lower",0)?LOWER_SEMANTICS:UPPER_SEMANTICS);
                for(HEnclList::const_iterator r=re.first.begin(); r!=re.first.end(); ++r) {
                    fprintf(out,"r %s",r->first.name().c_str()); dump_box(out,r->second.bounding_box()); fprintf(out,"\n"); }
                for(HEnclList::const_iterator r=re.second.begin(); r!=re.second.end(); ++r) {
                    fprintf(out,"f %s",r->first.name().c_str()); dump_box(out,r->second.bounding_box()); fprintf(out,"\n"); }
                if(geti(kv,"disc",0)) {
                    // what UpperReachEvolveWorker adjoins for this enclosure (include/workers.h:144-148); disc=2: only the union
                    if(geti(kv,"disc",0)==1) {
                    dump_set(out,"dreach",outer_approximation(re.first,grid,c.acc));
                    dump_set(out,"dfinal",outer_approximation(re.second,grid,c.acc)); }
                    ureach.adjoin(outer_approximation(re.first,grid,c.acc)); ufinal.adjoin(outer_approximation(re.second,grid,c.acc));
                }
            } catch(std::exception& e) {
                fprintf(out,"x %s\n",e.what());
            }
        }
        if(geti(kv,"disc",0)) { fprintf(out,"union\n"); dump_set(out,"ureach",ureach); dump_set(out,"ufinal",ufinal); }
        return 0;
    }

    if(mode=="worker") {
        // The batch seam itself: UpperReachEvolveWorker::get_result() (include/workers.h:51-81,102-153) over a strided
        // sample of the minced first-lap frontier, with boost::thread::hardware_concurrency() worker threads
        // (ORACLE_THREADS overrides), including the mutex-protected discretisation. This is the CPU baseline of bench.py.
        ReachabilityRestriction restriction(c.domain,grid,c.acc);
        HybridDenotableSet initial=restriction.outer_intersection_with(c.initial);
        if(geti(kv,"mince",1)) { for(HybridDenotableSet::locations_iterator li=initial.locations_begin(); li!=initial.locations_end(); ++li) li->second.mince(c.acc); }
        std::list<HEncl> all=to_enclosures(initial);
        int ncells=geti(kv,"ncells",64); int stride=geti(kv,"stride",1); int offset=geti(kv,"offset",0);
        std::list<HEncl> sample; int k=0;
        for(std::list<HEncl>::const_iterator it=all.begin(); it!=all.end() && (int)sample.size()<ncells; ++it,++k) { if(k<offset || (k-offset)%stride) continue; sample.push_back(*it); }
        unsigned long n=sample.size();
        boost::shared_ptr<EvolverInterface<HybridAutomatonInterface,HEncl> > evp(new ImageSetHybridEvolver(c.sys));
        { ReachabilityRestriction rr(c.domain,grid,c.acc); evp->tune_settings(rr.bounding_box(),grid,c.acc); }
        HybridTime ht(T,c.chain?1:c.events);
        bool ignore_act=c.chain; ContinuousEvolutionDirection dir=DIRECTION_FORWARD;
        std::chrono::steady_clock::time_point w0=std::chrono::steady_clock::now();
        UpperReachEvolveWorker worker(evp,sample,ht,grid,c.acc,ignore_act,dir);
        std::pair<HybridDenotableSet,HybridDenotableSet> res=worker.get_result();
        double ws=std::chrono::duration<double>(std::chrono::steady_clock::now()-w0).count();
        fprintf(out,"worker seconds %.6f cells %lu frontier %lu reach_cells %lu evolve_cells %lu threads %u\n",ws,n,(unsigned long)all.size(),
                (unsigned long)res.first.size(),(unsigned long)res.second.size(),boost::thread::hardware_concurrency());
        if(geti(kv,"dump",0)) { dump_set(out,"reach",res.first); dump_set(out,"evolve",res.second); }   // full result, for fingerprints
        return 0;
    }

    if(mode=="epslower") {
        // epsilon_lower_chain_reach (src/reachability_analyser.cpp:497-627) with no constraint set. prune=0 switches
        // enable_lower_pruning off (order-independent result); prune=1 keeps the default, whose rand() draws follow the
        // order in which LowerReachEpsilonWorker's threads finish -- run it with ORACLE_THREADS=1 for a reproducible golden.
        analyser.settings().enable_lower_pruning=geti(kv,"prune",1)!=0;
        std::pair<HybridDenotableSet,HybridFloatVector> r; std::string err;
        std::chrono::steady_clock::time_point t0=std::chrono::steady_clock::now();
        try { r=analyser.epsilon_lower_chain_reach(c.initial); } catch(std::exception& e) { err=e.what(); }
        double secs=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
        fprintf(out,"seconds %.3f err '%s' prune %d threads %u\n",secs,err.c_str(),geti(kv,"prune",1),boost::thread::hardware_concurrency());
        if(!c.cons_text.empty()) { std::stringstream ss(c.cons_text); std::string line; while(std::getline(ss,line)) fprintf(out,"cons| %s\n",line.c_str()); }
        for(HybridFloatVector::const_iterator ei=r.second.begin(); ei!=r.second.end(); ++ei) {
            fprintf(out,"epsilon loc %s",ei->first.name().c_str());
            for(uint i=0;i<ei->second.size();++i) fprintf(out," %a",ei->second[i]);
            fprintf(out,"\n");
        }
        dump_set(out,"reach",r.first);
        return 0;
    }

    std::ofstream logf;
    std::string logpath=kv.count("log")?kv["log"]:std::string("/tmp/ref_harness.log");
    if(mode=="time") { analyser.verbosity=4; redirect_log(logpath.c_str()); }
    std::chrono::steady_clock::time_point t0=std::chrono::steady_clock::now();
    HybridDenotableSet reach, evolve;
    std::string err;
    try {
        if(geti(kv,"lower",0)) { make_lpair(reach,evolve)=analyser.lower_reach_evolve(c.initial,HybridTime(T,c.events)); }
        else if(c.chain) { reach=analyser.outer_chain_reach(c.initial,geti(kv,"backward",0)?DIRECTION_BACKWARD:DIRECTION_FORWARD); }
        else { make_lpair(reach,evolve)=analyser.upper_reach_evolve(c.initial,HybridTime(T,c.events)); }
    } catch(std::exception& e) { err=e.what(); }
    std::chrono::steady_clock::time_point t1=std::chrono::steady_clock::now();
    double secs=std::chrono::duration<double>(t1-t0).count();
    if(mode=="time") {
        std::clog.flush();
        // cells evolved = sum of "Initial size" (src/reachability_analyser.cpp:190) + serial first lap "Starting from" (:347)
        std::ifstream lf(logpath.c_str()); std::string line; unsigned long evolved=0;
        while(std::getline(lf,line)) {
            size_t p=line.find("Initial size = "); if(p!=std::string::npos) evolved+=strtoul(line.c_str()+p+15,0,10);
            p=line.find("Starting from "); if(p!=std::string::npos) evolved+=strtoul(line.c_str()+p+14,0,10);
        }
        fprintf(out,"time seconds %.6f cells_evolved %lu reach_cells %lu threads %u err '%s'\n",secs,evolved,(unsigned long)reach.size(),boost::thread::hardware_concurrency(),err.c_str());
        return 0;
    }
    fprintf(out,"seconds %.3f err '%s'\n",secs,err.c_str());
    if(!c.cons_text.empty()) {
        // the initial cells of the constraint set (ReachabilityRestriction::outer_intersection_with) and its text form
        std::stringstream ss(c.cons_text); std::string line;
        while(std::getline(ss,line)) fprintf(out,"cons| %s\n",line.c_str());
        ReachabilityRestriction rr(c.domain,grid,c.acc);
        dump_set(out,"initial",rr.outer_intersection_with(c.initial));
    }
    dump_set(out,"reach",reach);
    dump_set(out,"evolve",evolve);
    return 0;
}
