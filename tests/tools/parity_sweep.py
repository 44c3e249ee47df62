"""Randomised parity sweep (development aid, CPU only): random initial boxes / accuracies / horizons for van der Pol, the
bouncing ball and the watertank, each run through the compiled reference (oracle/_ref/ref_harness, one worker thread) and
through the host build of the kernel headers + analyser mirror (tests/tools/hostchain.cpp); cell sets and epsilon vectors must
be identical bit for bit, and a run the reference rejects (domain check) must be rejected too.
  g++ -std=c++17 -O1 -frounding-math -w -x c++ -o build/hostchain_eps tests/tools/hostchain.cpp \
      release-1.0_b200/csrc/ab200_system.cpp release-1.0_b200/csrc/ab200_paving.cpp
  python tests/tools/parity_sweep.py <seed> <seconds>        (scratch files under /tmp/fz)"""
import random, subprocess, os, sys, time
R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.makedirs('/tmp/fz', exist_ok=True); H=R+'/oracle/_ref/ref_harness'; HC=R+'/build/hostchain_eps'
random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 1)
env=dict(os.environ, ORACLE_THREADS='1')
t_end=time.time()+float(sys.argv[2]) if len(sys.argv)>2 else time.time()+600
n=0; bad=0
def box(lo,hi,wmax):
    a=random.uniform(lo,hi-0.01); w=random.choice([0.0, random.uniform(0.001,wmax)]); b=min(hi,a+w); return a,b
while time.time()<t_end:
    kinds=['vdp_upper','vdp_lower','bb_chain','bb_lower','vdp_eps','bb_eps','wt_upper']
    if os.environ.get('SWEEP_CONS'): kinds=['vdp_cons_upper','vdp_cons_lower','vdp_cons_eps']      # constraint-set initial sets (disc)
    if os.environ.get('SWEEP_WT'): kinds=['wt_lower','wt_eps','wt_upper']          # watertank: transitions, sqrt dynamics, interval parameter
    if os.environ.get('SWEEP_SPRINGS'): kinds=['sp_upper','sp_eps']              # five variables: two-word keys (build/hostchain2_eps, -DAB_KW=2)
    if os.environ.get('SWEEP_MISC'): kinds=['bb_bchain','pend_upper','boost_upper']   # backward chain reach, sin dynamics, three variables (van der Pol backward chains take the reference minutes)
    kind=random.choice(kinds)
    out='/tmp/fz/g%s_%d.txt'%(sys.argv[1],n)
    if kind=='bb_bchain':
        acc=random.choice([3,4,5]); v=box(-3.0,4.0,0.5); x=box(0.0,1.5,0.3)
        init='%r,%r,%r,%r'%(v[0],v[1],x[0],x[1]); dom='-10.1,10.1,-0.1,2.1'; sysf='bouncingball'
        ra=['bouncingball','reach','acc=%d'%acc,'backward=1','init='+init]; ha=['bchain','acc=%d'%acc]
    elif kind=='pend_upper':
        acc=random.choice([3,4]); x=box(-2.0,2.0,0.3); y=box(-1.5,1.5,0.3); T=random.choice([0.5,1.0])
        init='%r,%r,%r,%r'%(x[0],x[1],y[0],y[1]); dom='-3,3,-3,3'; sysf='pendulum'
        ra=['pendulum','reach','acc=%d'%acc,'T=%r'%T,'init='+init]; ha=['upper','acc=%d'%acc,'T=%r'%T]
    elif kind=='boost_upper':
        acc=random.choice([3,4]); a=box(0.5,2.0,0.1); b=box(0.5,3.0,0.1); T=random.choice([0.003,0.011])
        init='0.0,0.0,%r,%r,%r,%r'%(a[0],a[1],b[0],b[1]); dom='0.0,0.001,0.0,3.0,0.0,6.0'; sysf='boost'
        ra=['boost','reach','acc=%d'%acc,'chain=0','T=%r'%T,'init='+init]; ha=['upper','acc=%d'%acc,'T=%r'%T,'loc=1']
    elif kind=='vdp_bupper':
        acc=random.choice([3,4]); x=box(-2.0,2.0,0.3); y=box(-2.5,2.5,0.3); dom='-3,3,-3,3'; sysf='vanderpol'
        init='%r,%r,%r,%r'%(x[0],x[1],y[0],y[1])
        ra=['vanderpol','reach','chain=1','backward=1','acc=%d'%acc,'init='+init]; ha=['bchain','acc=%d'%acc]
    elif kind.startswith('vdp_cons'):
        acc=random.choice([3,4,5]); cx=random.uniform(-1.5,1.5); cy=random.uniform(cx,2.0); cr=random.uniform(0.05,0.4)
        big=random.choice([0,1]); w=cr*random.uniform(0.8,1.6)
        init=None if big else '%r,%r,%r,%r'%(cx-w,cx+w,cy-w,cy+w)
        dom='-3,3,-3,3'; sysf='vanderpol'
        cons=['cons=disc','cc=%r,%r'%(cx,cy),'cr=%r'%cr]+([] if big else ['init='+init])
        if os.environ.get('SWEEP_CONS')=='band':      # two functions (x0+x1, x0*x1) in a random codomain box
            sm=cx+cy; pm=cx*cy; ws=random.uniform(0.05,0.5); wp=random.uniform(0.05,0.5)
            cons=['cons=band','cb=%r,%r,%r,%r'%(sm-ws,sm+ws,pm-wp,pm+wp)]+([] if big else ['init='+init])
        if kind=='vdp_cons_upper': T=random.choice([0.3,0.7]); ra=['vanderpol','reach','acc=%d'%acc,'T=%r'%T]+cons; ha=['upper','acc=%d'%acc,'T=%r'%T]
        elif kind=='vdp_cons_lower': T=random.choice([0.7,1.5]); ra=['vanderpol','reach','lower=1','acc=%d'%acc,'T=%r'%T]+cons; ha=['lower','acc=%d'%acc,'T=%r'%T]
        else: pr=random.choice([0,1]); ra=['vanderpol','epslower','acc=%d'%acc,'prune=%d'%pr]+cons; ha=['epslower','acc=%d'%acc,'prune=%d'%pr]
        init=None
    elif kind.startswith('vdp'):
        acc=random.choice([5,6] if os.environ.get('SWEEP_DEEP') else [2,3,4]); x=box(-2.0,2.0,0.05 if os.environ.get('SWEEP_DEEP') else 0.4); y=box(-2.5,2.5,0.05 if os.environ.get('SWEEP_DEEP') else 0.4)
        init='%r,%r,%r,%r'%(x[0],x[1],y[0],y[1]); dom='-3,3,-3,3'; sysf='vanderpol'
        if kind=='vdp_upper': T=random.choice([0.7,2.0,3.1]); ra=['vanderpol','reach','acc=%d'%acc,'init='+init,'T=%r'%T]; ha=['upper','acc=%d'%acc,'T=%r'%T]
        elif kind=='vdp_lower': T=random.choice([0.7,1.5]); ra=['vanderpol','reach','lower=1','acc=%d'%acc,'init='+init,'T=%r'%T]; ha=['lower','acc=%d'%acc,'T=%r'%T]
        else: pr=random.choice([0,1]); ra=['vanderpol','epslower','acc=%d'%acc,'init='+init,'prune=%d'%pr]; ha=['epslower','acc=%d'%acc,'prune=%d'%pr]
    elif kind.startswith('bb'):
        acc=random.choice([6,7] if os.environ.get('SWEEP_DEEP') else [3,4,5]); v=box(-3.0,3.0,0.1 if os.environ.get('SWEEP_DEEP') else 0.5); x=box(0.0,2.0,0.1 if os.environ.get('SWEEP_DEEP') else 0.5)
        init='%r,%r,%r,%r'%(v[0],v[1],x[0],x[1]); dom='-10.1,10.1,-0.1,2.1'; sysf='bouncingball'
        if kind=='bb_chain': ra=['bouncingball','reach','acc=%d'%acc,'init='+init]; ha=['chain','acc=%d'%acc]
        elif kind=='bb_lower': T=random.choice([0.5,1.5]); ra=['bouncingball','reach','chain=0','lower=1','acc=%d'%acc,'init='+init,'T=%r'%T]; ha=['lower','acc=%d'%acc,'T=%r'%T]
        else: pr=random.choice([0,1]); ra=['bouncingball','epslower','acc=%d'%acc,'init='+init,'prune=%d'%pr]; ha=['epslower','acc=%d'%acc,'prune=%d'%pr]
    elif kind.startswith('sp_'):
        acc=2; sysf='springs'; dom='0.0,40.0,-2.0,2.0,-3.0,3.0,-1.0,3.0,0.0,4.0'
        a=box(-0.3,0.3,0.1); b=box(-0.3,0.3,0.1); c=box(0.0,0.8,0.1); e=box(2.5,3.5,0.1)
        init='0,0,%r,%r,%r,%r,%r,%r,%r,%r'%(a[0],a[1],b[0],b[1],c[0],c[1],e[0],e[1])
        if kind=='sp_upper': T=random.choice([0.5,1.0,1.5]); ra=['springs','reach','acc=2','T=%r'%T,'init='+init]; ha=['upper','acc=2','T=%r'%T]
        else: L=random.choice([0.5,1.0]); pr=random.choice([0,1]); ra=['springs','epslower','acc=2','prune=%d'%pr,'lock=%r'%L,'init='+init]; ha=['epslower','acc=2','prune=%d'%pr,'lock=%r'%L]
    elif kind in ('wt_lower','wt_eps'):
        acc=random.choice([4] if os.environ.get('SWEEP_DEEP') else [2,3]); x=box(5.5,8.0,0.05 if os.environ.get('SWEEP_DEEP') else 0.3); sysf='watertank'; dom='0.0,1.0,4.5,9.0'; init=None
        ini=['initr=1.0,1.0,%r,%r'%x,'initf=0.0,0.0,%r,%r'%x]; hini=['loc=2','init=1.0,1.0,%r,%r'%x,'loc=1','init=0.0,0.0,%r,%r'%x]
        if kind=='wt_lower': T=random.choice([0.5,2.0,6.0]); ra=['watertank','reach','lower=1','acc=%d'%acc,'T=%r'%T]+ini; ha=['lower','acc=%d'%acc,'T=%r'%T]+hini
        else: pr=random.choice([0,1,1]); ra=['watertank','epslower','acc=%d'%acc,'prune=%d'%pr]+ini; ha=['epslower','acc=%d'%acc,'prune=%d'%pr,'subsys='+R+'/tests/golden/watertank_midparams.sys']+hini
    else:
        acc=random.choice([2,3]); T=random.choice([0.5,1.0]); x=box(5.5,8.0,0.3)
        ra=['watertank','reach','acc=%d'%acc,'T=%r'%T,'initr=1.0,1.0,%r,%r'%x,'initf=0.0,0.0,%r,%r'%x]; sysf='watertank'
        ha=['upper','acc=%d'%acc,'T=%r'%T,'loc=2','init=1.0,1.0,%r,%r'%x,'loc=1','init=0.0,0.0,%r,%r'%x]; dom='0.0,1.0,4.5,9.0'; init=None
    try:
        subprocess.run([H]+ra+['out='+out],env=env,timeout=60,check=True,stdout=subprocess.DEVNULL,stderr=subprocess.DEVNULL)
    except Exception as e:
        print('ref skip',kind,ra[2:],type(e).__name__,flush=True); continue
    txt=open(out).read()
    referr = "err ''" not in txt
    args=[(HC.replace('hostchain_eps','hostchain2_eps') if sysf in ('springs','boost') else HC),R+'/tests/golden/%s.sys'%sysf,out]+ha+['domain='+dom]+(['init='+init] if init else [])
    try:
        o=subprocess.run(args,timeout=240,stdout=subprocess.PIPE,stderr=subprocess.STDOUT).stdout.decode()
    except subprocess.TimeoutExpired:
        print('host timeout',kind,ra[2:],flush=True); continue
    ok='cell sets identical' in o and 'differs' not in o
    if referr:
        ok = 'rc=' in o      # reference raised: the host build must fail too
    n+=1
    if not ok:
        bad+=1; print('MISMATCH',kind,' '.join(ra),'| referr' if referr else '', o[-300:].replace('\n',' / '),flush=True)
        os.rename(out,'/tmp/fz/bad%s_%d.txt'%(sys.argv[1],n))
    else: print('ok',kind,ra[2:],'referr' if referr else '',flush=True)
print('done',n,'cases',bad,'bad')
