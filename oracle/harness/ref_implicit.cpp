// Oracle tooling only. Reads a TaylorModel in the text form printed by tests/tools (header "name as=.. n=.. e=.."
// then "[e0 e1 ..] coef" lines), calls the reference's implicit() / jacobian2_range on it and prints the result.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include "ariadne.h"
#include "taylor_model.h"
using namespace Ariadne;
namespace Ariadne { Interval jacobian2_range(const TaylorModel& tm); }
static void dump(const char* name, const TaylorModel& t) {
    printf("%s as=%u n=%u e=%a sweep=%g deg=%u\n", name, t.argument_size(), (unsigned)t.number_of_nonzeros(), t.error(), t.sweep_threshold(), t.maximum_degree());
    for(TaylorModel::const_iterator it=t.begin(); it!=t.end(); ++it) {
        printf("  ["); for(uint j=0;j<t.argument_size();++j) printf("%u ",(unsigned)it->key()[j]); printf("] %a\n", (double)it->data()); }
}
int main() {
    std::string line; std::getline(std::cin,line);
    unsigned as=0,n=0; char name[64]; char es[64];
    sscanf(line.c_str(),"%63s as=%u n=%u e=%63s",name,&as,&n,es);
    TaylorModel f(as);
    for(unsigned k=0;k<n;++k) {
        std::getline(std::cin,line);
        size_t a=line.find('['), b=line.find(']');
        std::istringstream es2(line.substr(a+1,b-a-1)); MultiIndex mi(as); for(unsigned j=0;j<as;++j) { unsigned e; es2>>e; mi[j]=e; }
        f.expansion().append(mi,strtod(line.c_str()+b+1,0));
    }
    f.set_error(strtod(es,0));
    dump("f",f);
    Interval J=jacobian2_range(f);
    printf("jacobian2_range %a %a\n",J.lower(),J.upper());
    try { TaylorModel h=implicit(f); dump("h",h); }
    catch(std::exception& e) { printf("exception %s\n",e.what()); }
    return 0;
}
