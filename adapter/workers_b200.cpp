// adapter/workers_b200.cpp -- see workers_b200.h.
#include "ariadne.h"
#include "workers.h"
#include "workers_b200.h"
#include "evolver_b200.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace Ariadne {

namespace {

void emit_tape(std::ostream& f, const RealExpression& e, const std::map<std::string,int>& var_index) {
    Operator op=e.op();
    char buf[96];
    if(op==VAR) {
        std::stringstream ss; ss<<e; std::string name=ss.str();
        std::map<std::string,int>::const_iterator it=var_index.find(name);
        if(it==var_index.end()) throw std::runtime_error("b200_system_text: unknown variable '"+name+"'");
        f<<" V"<<it->second; return;
    }
    if(op==CNST || op==PAR) {
        Interval v=evaluate(e,Vector<Interval>(1));
        snprintf(buf,sizeof(buf)," C%a,%a",v.lower(),v.upper()); f<<buf; return;
    }
    List<RealExpression> sub=e.subexpressions();
    for(uint i=0;i<sub.size();++i) emit_tape(f,sub[i],var_index);
    const char* name=0;
    switch(op) {
        case ADD: name="ADD"; break; case SUB: name="SUB"; break; case MUL: name="MUL"; break; case DIV: name="DIV"; break;
        case NEG: name="NEG"; break; case POS: name="POS"; break; case REC: name="REC"; break; case SQR: name="SQR"; break;
        case SQRT: name="SQRT"; break; case EXP: name="EXP"; break; case LOG: name="LOG"; break; case SIN: name="SIN"; break;
        case COS: name="COS"; break; case TAN: name="TAN"; break; case ABS: name="ABS"; break; case MAX: name="MAX"; break; case MIN: name="MIN"; break;
        default: throw std::runtime_error("b200_system_text: unsupported operator");
    }
    f<<" "<<name;
}

// lie_derivative(g,dynamic) as ScalarFunction arithmetic builds it (src/function.cpp:1308-1331,:863-909): r=dg/dx0*f0;
// r=r+dg/dxi*fi with Ariadne::derivative on the expressions. Needed by the lower-semantics discard test
// (src/hybrid_evolver-image.cpp:612-632) and the grid-side jump test (src/set_checker.cpp:148-161).
void emit_lie(std::ostream& f, const RealExpression& g, const std::map<RealVariable,RealExpression>& dyn, const RealSpace& cs,
              const std::map<std::string,int>& var_index) {
    std::map<RealVariable,RealExpression>::const_iterator d=dyn.begin();
    try {
        RealExpression lie=derivative(g,cs.variable(0))*d->second; ++d;
        for(uint i=1;i<cs.size();++i,++d) lie=lie+derivative(g,cs.variable(i))*d->second;
        f<<"    lie"; emit_tape(f,lie,var_index); f<<" ;\n";
    } catch(std::exception&) {
        // the reference itself cannot differentiate this function (e.g. min(-iL,iL), src/function.cpp:1329)
        f<<"    # lie derivative not computable in the reference: derivative() throws\n";
    }
}

} // namespace

std::string b200_system_text(const HybridIOAutomaton& sys) {
    std::ostringstream f;
    HybridSpace space=sys.state_space();
    std::map<std::string,int> loc_index; int L=0;
    for(HybridSpace::const_iterator it=space.begin(); it!=space.end(); ++it) loc_index[it->first.name()]=L++;
    f<<"system "<<sys.name()<<"\nlocations "<<L<<"\n";
    for(HybridSpace::const_iterator it=space.begin(); it!=space.end(); ++it) {
        DiscreteLocation loc=it->first;
        RealSpace cs=sys.continuous_state_space(loc);
        std::map<std::string,int> var_index;
        for(uint i=0;i<cs.size();++i) var_index[cs.variable(i).name()]=i;
        Set<DiscreteEvent> events=sys.events(loc);
        f<<"location "<<loc_index[loc.name()]<<" "<<loc.name()<<" n "<<cs.size()<<" events "<<events.size()<<"\n";
        const std::map<RealVariable,RealExpression>* dyn=0;
        for(std::list<DiscreteIOMode>::const_iterator m=sys.modes().begin(); m!=sys.modes().end(); ++m) {
            if(!(m->location()==loc)) continue;
            dyn=&m->dynamics();
            uint i=0;
            for(std::map<RealVariable,RealExpression>::const_iterator d=m->dynamics().begin(); d!=m->dynamics().end(); ++d,++i) {
                f<<"  dyn "<<i; emit_tape(f,d->second,var_index); f<<" ;\n"; }
        }
        int k=0;
        for(Set<DiscreteEvent>::const_iterator e=events.begin(); e!=events.end(); ++e,++k) {
            EventKind kind=sys.event_kind(loc,*e);
            const char* kn = kind==INVARIANT?"invariant":kind==URGENT?"urgent":"permissive";
            int target = kind==INVARIANT ? loc_index[loc.name()] : loc_index[sys.target(loc,*e).name()];
            f<<"  event "<<k<<" "<<e->name()<<" kind "<<kn<<" target "<<target<<" id "<<k<<"\n";
            if(kind==INVARIANT) {
                for(std::list<DiscreteIOMode>::const_iterator m=sys.modes().begin(); m!=sys.modes().end(); ++m) {
                    if(!(m->location()==loc)) continue;
                    std::map<DiscreteEvent,RealExpression>::const_iterator inv=m->invariants().find(*e);
                    f<<"    guard"; emit_tape(f,inv->second,var_index); f<<" ;\n";
                    emit_lie(f,inv->second,m->dynamics(),cs,var_index);
                }
            } else {
                for(std::list<DiscreteIOTransition>::const_iterator t=sys.transitions().begin(); t!=sys.transitions().end(); ++t) {
                    if(!(t->source()==loc && t->event()==*e)) continue;
                    f<<"    guard"; emit_tape(f,t->activation(),var_index); f<<" ;\n";
                    emit_lie(f,t->activation(),*dyn,cs,var_index);
                    uint i=0;
                    for(std::map<RealVariable,RealExpression>::const_iterator r=t->reset().begin(); r!=t->reset().end(); ++r,++i) {
                        f<<"    reset "<<i; emit_tape(f,r->second,var_index); f<<" ;\n"; }
                    break;
                }
            }
        }
    }
    return f.str();
}

std::string b200_constraint_text(const HybridIOAutomaton& sys, const DiscreteLocation& loc, const Box& domain,
                                 const List<RealExpression>& functions, const Box& codomain) {
    std::ostringstream f; char buf[64];
    RealSpace cs=sys.continuous_state_space(loc);
    std::map<std::string,int> var_index;
    for(uint i=0;i<cs.size();++i) var_index[cs.variable(i).name()]=i;
    f<<"set "<<loc.name()<<" m "<<functions.size()<<"\n  domain";
    for(uint i=0;i<domain.size();++i) { snprintf(buf,sizeof(buf)," %a %a",domain[i].lower(),domain[i].upper()); f<<buf; }
    f<<"\n  codomain";
    for(uint i=0;i<codomain.size();++i) { snprintf(buf,sizeof(buf)," %a %a",codomain[i].lower(),codomain[i].upper()); f<<buf; }
    f<<"\n";
    for(uint i=0;i<functions.size();++i) { f<<"  f "<<i; emit_tape(f,functions[i],var_index); f<<" ;\n"; }
    return f.str();
}

namespace {

// One device context per automaton (by name), created on first use; lives for the process.
struct B200Context {
    ab200_ctx* ctx;
    std::string system_name, system_text;
    B200Context() : ctx(0) { }
};
B200Context& b200_context(const HybridIOAutomaton& sys) {
    static std::map<std::string,B200Context> cache;
    B200Context& c=cache[sys.name()];
    if(!c.ctx) {
        int device=0; if(const char* d=getenv("AB200_DEVICE")) device=atoi(d);
        if(ab200_ctx_create(&c.ctx,device)!=AB200_OK) throw std::runtime_error("ab200_ctx_create: no usable sm_100 GPU");
        // device sizing: generous per-cell scratch and box lists (overridable: AB200_ARENA_WORDS, AB200_MAX_BOXES)
        ab200_settings st; ab200_default_settings(&st);
        st.arena_words=262144; st.max_boxes=512;
        if(const char* e=getenv("AB200_ARENA_WORDS")) st.arena_words=atoi(e);
        if(const char* e=getenv("AB200_MAX_BOXES")) st.max_boxes=atoi(e);
        if(ab200_set_settings(c.ctx,&st)!=AB200_OK) throw std::runtime_error(std::string("ab200_set_settings: ")+ab200_last_error(c.ctx));
        c.system_name=sys.name();
    }
    // the automaton may have had parameters substituted since the last call (epsilon_lower_chain_reach evolves the system
    // with its interval parameters at their midpoints, src/reachability_analyser.cpp:517-527): upload the text when it changed
    std::string text=b200_system_text(sys);
    if(text!=c.system_text) {
        if(ab200_set_system_text(c.ctx,text.c_str())!=AB200_OK) throw std::runtime_error(std::string("ab200_set_system_text: ")+ab200_last_error(c.ctx));
        c.system_text=text;
    }
    return c;
}

// The frontier enclosures are TaylorSet(cell box) (src/reachability_analyser.cpp:1244-1253): model i is
// TaylorModel::scaling(n,i,[a,b]). The library takes the cell box [a,b] itself; bounding_box() of the enclosure is
// range(scaling), a few ulps wider. The exact box is recovered by trying the doubles next to the range's end points and
// keeping the pair whose scaling is the enclosure's model (value, gradient and error all equal).
bool same_model(const TaylorModel& a, const TaylorModel& b) {
    return a.argument_size()==b.argument_size() && a.expansion()==b.expansion() && a.error()==b.error();
}
Interval recover_cell_interval(const TaylorSet& ts, uint i) {
    const uint n=ts.dimension();
    const TaylorModel& m=ts.models()[i];
    Interval r=m.range();
    double lo=r.lower(), hi=r.upper();
    std::vector<double> los, his;
    { double x=lo; for(int k=0;k<8;++k) { los.push_back(x); x=nextafter(x,1e300); } }
    { double x=hi; for(int k=0;k<8;++k) { his.push_back(x); x=nextafter(x,-1e300); } }
    for(size_t p=0;p<los.size();++p) for(size_t q=0;q<his.size();++q) {
        if(!(los[p]<=his[q])) continue;
        TaylorModel s=TaylorModel::scaling(n,i,Interval(los[p],his[q]));
        if(same_model(s,m)) return Interval(los[p],his[q]);
    }
    throw std::runtime_error("B200UpperReachEvolveWorker: the enclosure is not a grid-cell box (only TaylorSet(Box) enclosures are supported)");
}

} // namespace

B200UpperReachEvolveWorker::B200UpperReachEvolveWorker(const EvolverType& evolver, std::list<EnclosureType>& initial_enclosures,
        const HybridTime& time, const HybridGrid& grid, const int& accuracy, const bool& ignore_activations,
        const ContinuousEvolutionDirection& continuous_direction)
    : _evolver(evolver), _initial_enclosures(initial_enclosures), _time(time), _grid(grid), _accuracy(accuracy),
      _ignore_activations(ignore_activations), _continuous_direction(continuous_direction) { }

std::pair<HybridDenotableSet,HybridDenotableSet> B200UpperReachEvolveWorker::get_result() {
    const HybridIOAutomaton* sys=dynamic_cast<const HybridIOAutomaton*>(&_evolver->system());
    const ImageSetHybridEvolver* ise=dynamic_cast<const ImageSetHybridEvolver*>(_evolver.get());
    if(!sys || !ise) throw std::runtime_error("B200UpperReachEvolveWorker: needs an ImageSetHybridEvolver over a HybridIOAutomaton");
    B200Context& bc=b200_context(*sys);
    ab200_ctx* ctx=bc.ctx;
    // ImageSetHybridEvolver::tune_settings has been applied to the evolver by _get_tuned_evolver; hand the result over
    const std::map<DiscreteLocation,Float>& hmax=ise->settings().maximum_step_size();
    const HybridFloatVector& refw=ise->settings().reference_enclosure_widths();
    for(HybridGrid::locations_const_iterator gi=_grid.locations_begin(); gi!=_grid.locations_end(); ++gi) {
        const DiscreteLocation& q=gi->first; const Grid& g=gi->second;
        int li=ab200_location_index(ctx,q.name().c_str());
        if(li<0) throw std::runtime_error("B200UpperReachEvolveWorker: unknown location "+q.name());
        double rw[AB200_MAX_DIM]={0}, org[AB200_MAX_DIM]={0}, len[AB200_MAX_DIM]={0};
        const Vector<Float>& w=refw.find(q)->second;
        for(uint i=0;i<g.dimension();++i) { rw[i]=w[i]; org[i]=g.origin()[i]; len[i]=g.lengths()[i]; }
        if(ab200_set_location_tuning(ctx,li,hmax.find(q)->second,rw,org,len)!=AB200_OK) throw std::runtime_error(ab200_last_error(ctx));
    }
    if(ab200_set_accuracy(ctx,_accuracy)!=AB200_OK) throw std::runtime_error(ab200_last_error(ctx));

    std::vector<int32_t> loc; std::vector<double> box;
    while(!_initial_enclosures.empty()) {                            // consumes the list, as include/workers.h:120-121 does
        const EnclosureType& e=_initial_enclosures.front();
        loc.push_back(ab200_location_index(ctx,e.first.name().c_str()));
        const TaylorSet& ts=e.second;
        size_t o=box.size(); box.resize(o+AB200_BOX_STRIDE,0.0);
        for(uint i=0;i<ts.dimension();++i) { Interval c=recover_cell_interval(ts,i); box[o+2*i]=c.lower(); box[o+2*i+1]=c.upper(); }
        _initial_enclosures.pop_front();
    }
    ab200_cellset *reach=0,*evolve=0; ab200_status st;
    int rc=ab200_upper_reach_evolve(ctx,loc.size(),loc.empty()?0:&loc[0],box.empty()?0:&box[0],_time.continuous_time(),
                                    _time.discrete_time(),_ignore_activations?1:0,_continuous_direction==DIRECTION_FORWARD?0:1,
                                    &reach,&evolve,&st);
    switch(rc) {                                                     // same exception kinds as include/workers.h:159-166
        case AB200_OK: break;
        case AB200_E_SQRT: throw SqrtNumericException(ab200_last_error(ctx));
        case AB200_E_TIMEOUT: throw TimeoutException();
        case AB200_E_WORKING_SETS: throw WorkingSetTooLargeException(ab200_last_error(ctx));
        default: throw std::runtime_error(std::string("libariadne_b200: ")+ab200_last_error(ctx));
    }
    std::pair<HDS,HDS> result(_to_denotable(reach),_to_denotable(evolve));
    ab200_cellset_free(reach); ab200_cellset_free(evolve);
    return result;
}

// ---- LowerReachEpsilonWorker ---------------------------------------------------------------------------------------------
B200LowerReachEpsilonWorker::B200LowerReachEpsilonWorker(const EvolverType& evolver, EL& initial_enclosures, const HybridTime& time,
        const HybridGrid& grid, const int& accuracy)
    : _evolver(evolver), _initial_enclosures(initial_enclosures), _time(time), _grid(grid), _accuracy(accuracy) { }

tuple<std::pair<B200LowerReachEpsilonWorker::HUM,B200LowerReachEpsilonWorker::HUM>,B200LowerReachEpsilonWorker::EL,HybridDenotableSet,HybridFloatVector>
B200LowerReachEpsilonWorker::get_result() {
    const HybridIOAutomaton* sys=dynamic_cast<const HybridIOAutomaton*>(&_evolver->system());
    const ImageSetHybridEvolver* ise=dynamic_cast<const ImageSetHybridEvolver*>(_evolver.get());
    if(!sys || !ise) throw std::runtime_error("B200LowerReachEpsilonWorker: needs an ImageSetHybridEvolver over a HybridIOAutomaton");
    ab200_ctx* ctx=b200_context(*sys).ctx;
    const std::map<DiscreteLocation,Float>& hmax=ise->settings().maximum_step_size();
    const HybridFloatVector& refw=ise->settings().reference_enclosure_widths();
    std::vector<DiscreteLocation> locs;                               // library order = map order of the grid
    for(HybridGrid::locations_const_iterator gi=_grid.locations_begin(); gi!=_grid.locations_end(); ++gi) {
        const DiscreteLocation& q=gi->first; const Grid& g=gi->second;
        int li=ab200_location_index(ctx,q.name().c_str());
        if(li!=(int)locs.size()) throw std::runtime_error("B200LowerReachEpsilonWorker: location order differs for "+q.name());
        locs.push_back(q);
        double rw[AB200_MAX_DIM]={0}, org[AB200_MAX_DIM]={0}, len[AB200_MAX_DIM]={0};
        const Vector<Float>& w=refw.find(q)->second;
        for(uint i=0;i<g.dimension();++i) { rw[i]=w[i]; org[i]=g.origin()[i]; len[i]=g.lengths()[i]; }
        if(ab200_set_location_tuning(ctx,li,hmax.find(q)->second,rw,org,len)!=AB200_OK) throw std::runtime_error(ab200_last_error(ctx));
    }
    if(ab200_set_accuracy(ctx,_accuracy)!=AB200_OK) throw std::runtime_error(ab200_last_error(ctx));

    // the initial TaylorSets as enclosure records (consumes the list, as include/workers.h:256-257 does)
    std::vector<int32_t> eloc, nargs, nterms; std::vector<double> errors, coefs; std::vector<uint8_t> exps;
    while(!_initial_enclosures.empty()) {
        const EnclosureType& e=_initial_enclosures.front();
        const TaylorSet& ts=e.second;
        eloc.push_back(ab200_location_index(ctx,e.first.name().c_str()));
        nargs.push_back((int32_t)ts.argument_size());
        for(uint i=0;i<ts.dimension();++i) {
            const TaylorModel& m=ts.models()[i];
            nterms.push_back((int32_t)m.number_of_nonzeros()); errors.push_back(m.error());
            for(TaylorModel::const_iterator it=m.begin(); it!=m.end(); ++it) {
                for(uint j=0;j<m.argument_size();++j) exps.push_back((uint8_t)it->key()[j]);
                coefs.push_back(it->data());
            }
        }
        _initial_enclosures.pop_front();
    }
    const size_t n_units=eloc.size();
    ab200_enclosures* initial=0;
    if(ab200_enclosures_create(ctx,n_units,n_units?&eloc[0]:0,n_units?&nargs[0]:0,nterms.empty()?0:&nterms[0],errors.empty()?0:&errors[0],
                               exps.empty()?0:&exps[0],coefs.empty()?0:&coefs[0],&initial)!=AB200_OK)
        throw std::runtime_error(std::string("ab200_enclosures_create: ")+ab200_last_error(ctx));

    ab200_settings saved, low; ab200_ctx_get_settings(ctx,&saved);
    low=saved; low.semantics=1; low.maximum_enclosure_widths_ratio=ise->settings().maximum_enclosure_widths_ratio();
    ab200_boxlist *rb=0,*fb=0; ab200_enclosures* fe=0; ab200_status st;
    int rc=ab200_set_settings(ctx,&low);
    if(rc==AB200_OK) rc=ab200_reach_evolve_enclosures(ctx,n_units,0,0,initial,_time.continuous_time(),_time.discrete_time(),0,0,0,&rb,&fb,&fe,&st);
    ab200_set_settings(ctx,&saved);
    ab200_enclosures_free(initial);
    switch(rc) {                                                     // same exception kinds as include/workers.h:318-326
        case AB200_OK: break;
        case AB200_E_SQRT: throw SqrtNumericException(ab200_last_error(ctx));
        case AB200_E_TIMEOUT: throw TimeoutException();
        default: throw std::runtime_error(std::string("libariadne_b200: ")+ab200_last_error(ctx));
    }

    HDS reach(_grid), evolve_global(_grid);
    HUM adjoined, superposed; EL final_enclosures; HybridFloatVector epsilon;
    HybridSpace state_space=_evolver->system().state_space();
    for(HybridSpace::const_iterator hs=state_space.begin(); hs!=state_space.end(); ++hs)
        epsilon.insert(std::pair<DiscreteLocation,Vector<Float> >(hs->first,Vector<Float>(hs->second)));
    const uint64_t *roff=0,*foff=0,*uoff=0; const double *rbx=0,*fbx=0; const int32_t *rl=0,*fl=0,*el=0; uint64_t n_final=0,n_u=0;
    ab200_boxlist_view(rb,&roff,&rbx,&rl); ab200_boxlist_view(fb,&foff,&fbx,&fl);
    ab200_enclosures_view(fe,&n_final,&el,&uoff,&n_u);
    std::vector<uint8_t> ex; std::vector<double> co;
    for(size_t u=0;u<n_units;++u) {                                  // LowerReachEpsilonWorker::_compute, unit by unit
        HDS current_reach(_grid), current_evolve(_grid);
        for(uint64_t k=roff[u];k<roff[u+1];++k) {
            const DiscreteLocation& q=locs[rl[k]]; const uint n=state_space[q];
            Box b(n); for(uint i=0;i<n;++i) b[i]=Interval(rbx[k*AB200_BOX_STRIDE+2*i],rbx[k*AB200_BOX_STRIDE+2*i+1]);
            current_reach[q].adjoin_outer_approximation(b,_accuracy);
            epsilon[q]=max_elementwise(epsilon[q],b.widths());
        }
        for(uint64_t k=foff[u];k<foff[u+1];++k) {
            const DiscreteLocation& q=locs[fl[k]]; const uint n=state_space[q];
            Box b(n); for(uint i=0;i<n;++i) b[i]=Interval(fbx[k*AB200_BOX_STRIDE+2*i],fbx[k*AB200_BOX_STRIDE+2*i+1]);
            current_evolve[q].adjoin_outer_approximation(b,_accuracy);
        }
        reach.adjoin(current_reach); evolve_global.adjoin(current_evolve);
        for(HDS::locations_const_iterator it=current_evolve.locations_begin(); it!=current_evolve.locations_end(); ++it)
            superposed[it->first]+=it->second.size();
        // the final sets of the unit, location by location (ListSet keeps a map over locations), as TaylorSets
        for(size_t li=0;li<locs.size();++li) for(uint64_t k=uoff[u];k<uoff[u+1];++k) {
            if(el[k]!=(int32_t)li) continue;
            const uint n=state_space[locs[li]];
            Vector<TaylorModel> models(n);
            for(uint i=0;i<n;++i) {
                int32_t as=0,nt=0; double err=0;
                ab200_enclosures_model(fe,k,(int32_t)i,&as,&nt,&err,0,0);
                ex.assign((size_t)nt*as+1,0); co.assign((size_t)nt+1,0.0);
                ab200_enclosures_model(fe,k,(int32_t)i,&as,&nt,&err,&ex[0],&co[0]);
                Expansion<Float> f(as); f.reserve(nt);
                MultiIndex a(as);
                for(int32_t t=0;t<nt;++t) { for(int32_t j=0;j<as;++j) a.set(j,ex[(size_t)t*as+j]); f.append(a,co[t]); }
                models[i]=TaylorModel(f,err);
            }
            final_enclosures.push_back(EnclosureType(locs[li],TaylorSet(models)));
        }
    }
    ab200_boxlist_free(rb); ab200_boxlist_free(fb); ab200_enclosures_free(fe);
    for(HDS::locations_const_iterator it=evolve_global.locations_begin(); it!=evolve_global.locations_end(); ++it)
        adjoined[it->first]=it->second.size();
    Float accuracy_divider=(1<<_accuracy);
    HybridFloatVector grid_lengths=_grid.lengths();
    for(HybridSpace::const_iterator hs=state_space.begin(); hs!=state_space.end(); ++hs)
        epsilon[hs->first]+=grid_lengths[hs->first]/accuracy_divider;
    return make_tuple<std::pair<HUM,HUM>,EL,HDS,HybridFloatVector>(std::make_pair(adjoined,superposed),final_enclosures,reach,epsilon);
}

// ---- B2: EvolverInterface seam ------------------------------------------------------------------------------------------
void B200ImageSetHybridEvolver::_evolution(EnclosureListType& final, EnclosureListType& reachable, EnclosureListType& intermediate,
        const EnclosureType& initial, const TimeType& time, bool ignore_activations, ContinuousEvolutionDirection direction,
        Semantics semantics) const {
    const HybridIOAutomaton* sys=dynamic_cast<const HybridIOAutomaton*>(&this->system());
    if(!sys) throw std::runtime_error("B200ImageSetHybridEvolver: needs a HybridIOAutomaton");
    ab200_ctx* ctx=b200_context(*sys).ctx;
    // the settings of this object (tune_settings or the defaults) become the device tuning; boxes need no grid
    const std::map<DiscreteLocation,Float>& hmax=this->settings().maximum_step_size();
    const HybridFloatVector& refw=this->settings().reference_enclosure_widths();
    HybridSpace space=sys->state_space();
    for(HybridSpace::const_iterator it=space.begin(); it!=space.end(); ++it) {
        int li=ab200_location_index(ctx,it->first.name().c_str());
        double rw[AB200_MAX_DIM]={0}, org[AB200_MAX_DIM]={0}, len[AB200_MAX_DIM];
        for(int i=0;i<AB200_MAX_DIM;++i) len[i]=1.0;
        std::map<DiscreteLocation,Float>::const_iterator h=hmax.find(it->first);
        HybridFloatVector::const_iterator w=refw.find(it->first);
        if(h==hmax.end() || w==refw.end()) throw std::runtime_error("B200ImageSetHybridEvolver: no settings for location "+it->first.name());
        for(uint i=0;i<it->second;++i) rw[i]=w->second[i];
        if(ab200_set_location_tuning(ctx,li,h->second,rw,org,len)!=AB200_OK) throw std::runtime_error(ab200_last_error(ctx));
    }
    ab200_settings st; ab200_ctx_get_settings(ctx,&st);
    st.semantics = (semantics==LOWER_SEMANTICS) ? 1 : 0;
    st.maximum_enclosure_widths_ratio=this->settings().maximum_enclosure_widths_ratio();
    if(ab200_set_settings(ctx,&st)!=AB200_OK) throw std::runtime_error(ab200_last_error(ctx));

    const TaylorSet& ts=initial.second;
    int32_t loc=ab200_location_index(ctx,initial.first.name().c_str());
    double box[AB200_BOX_STRIDE]={0};
    for(uint i=0;i<ts.dimension();++i) {
        Interval c;
        try { c=recover_cell_interval(ts,i); }
        catch(std::runtime_error&) {
            if(semantics==LOWER_SEMANTICS) throw std::runtime_error("B200ImageSetHybridEvolver: LOWER_SEMANTICS needs a box enclosure");
            c=ts.models()[i].range();                                  // over-approximation: sound for UPPER_SEMANTICS
        }
        box[2*i]=c.lower(); box[2*i+1]=c.upper();
    }
    ab200_boxlist *rb=0,*fb=0; ab200_status status;
    int rc=ab200_reach_evolve_boxes(ctx,1,&loc,box,time.continuous_time(),time.discrete_time(),ignore_activations?1:0,
                                    direction==DIRECTION_FORWARD?0:1,&rb,&fb,&status);
    switch(rc) {
        case AB200_OK: break;
        case AB200_E_SQRT: throw SqrtNumericException(ab200_last_error(ctx));
        case AB200_E_TIMEOUT: throw TimeoutException();
        case AB200_E_WORKING_SETS: throw WorkingSetTooLargeException(ab200_last_error(ctx));
        default: throw std::runtime_error(std::string("libariadne_b200: ")+ab200_last_error(ctx));
    }
    std::vector<DiscreteLocation> locs;
    for(HybridSpace::const_iterator it=space.begin(); it!=space.end(); ++it) locs.push_back(it->first);   // library order
    ab200_boxlist* lists[2]={rb,fb}; EnclosureListType* outs[2]={&reachable,&final};
    for(int w=0;w<2;++w) {
        const uint64_t* off=0; const double* bx=0; const int32_t* bl=0;
        ab200_boxlist_view(lists[w],&off,&bx,&bl);
        for(uint64_t k=off[0];k<off[1];++k) {
            const DiscreteLocation& q=locs[bl[k]];
            const uint n=space[q];
            Box b(n);
            for(uint i=0;i<n;++i) b[i]=Interval(bx[k*AB200_BOX_STRIDE+2*i],bx[k*AB200_BOX_STRIDE+2*i+1]);
            outs[w]->adjoin(EnclosureType(q,TaylorSet(b)));
        }
        ab200_boxlist_free(lists[w]);
    }
    (void)intermediate;
}

// keys -> HybridDenotableSet. A key is (location, lattice index per coordinate at resolution 2^-accuracy of the base grid);
// the cell is adjoined through BDDTreeSet::adjoin_outer_approximation of its box shrunk by a quarter of its width on every
// side, which meets that cell and no other.
HybridDenotableSet B200UpperReachEvolveWorker::_to_denotable(const ab200_cellset* cs) const {
    HDS result(_grid);
    const uint64_t* keys=0; uint64_t count=0;
    ab200_cellset_view(cs,&keys,&count);
    std::vector<DiscreteLocation> locs;
    for(HybridGrid::locations_const_iterator gi=_grid.locations_begin(); gi!=_grid.locations_end(); ++gi) locs.push_back(gi->first);
    const HybridIOAutomaton* sys=dynamic_cast<const HybridIOAutomaton*>(&_evolver->system());
    ab200_ctx* ctx=b200_context(*sys).ctx;
    const double scale=std::ldexp(1.0,-_accuracy);
    for(uint64_t k=0;k<count;++k) {
        int32_t li=0; int64_t idx[AB200_MAX_DIM];
        ab200_key_decode(ctx,keys[k],&li,idx);
        // library location indices follow the HybridSpace order of b200_system_text, which is the grid's map order
        const DiscreteLocation& q=locs[li];
        const Grid& g=_grid[q];
        Box b(g.dimension());
        for(uint i=0;i<g.dimension();++i) {
            double w=g.lengths()[i]*scale, lo=g.origin()[i]+w*double(idx[i]);
            b[i]=Interval(lo+w/4,lo+3*w/4);
        }
        result[q].adjoin_outer_approximation(b,_accuracy);
    }
    return result;
}

} // namespace Ariadne
