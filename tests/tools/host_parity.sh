#!/bin/sh
# development aid: rebuilds the host tools and replays every golden through the host build of the kernel headers
# (bit-exact comparison of boxes and cell sets); run from the repository root
set -e
sh tests/tools/build.sh
g++ -std=c++17 -O1 -g -frounding-math -DAB_KW=2 -x c++ -o build/hostchain2 tests/tools/hostchain.cpp release-1.0_b200/csrc/ab200_system.cpp release-1.0_b200/csrc/ab200_paving.cpp
T="timeout 600"
zcat tests/golden/vanderpol_cells_acc6.txt.gz > build/vdp6.txt
zcat tests/golden/springs_reach_acc2.txt.gz > build/sp2.txt
$T build/hostdbg tests/golden/vanderpol.sys build/vdp6.txt | tail -1
$T build/hostdbg tests/golden/vanderpol.sys tests/golden/vanderpol_cells_acc3.txt | tail -1
$T build/hostdbg tests/golden/bouncingball.sys tests/golden/bouncingball_cells_acc4.txt | tail -1
$T build/hostdbg tests/golden/bouncingball.sys tests/golden/bouncingball_cells_acc5.txt | tail -1
$T build/hostlap tests/golden/vanderpol.sys tests/golden/vanderpol_reach_acc3.txt init=1.0,1.3,1.0,1.6 | tail -1
$T build/hostlap tests/golden/vanderpol.sys tests/golden/vanderpol_reach_acc4.txt init=-1.6,-1.3,0.3,0.8 | tail -1
$T build/hostchain tests/golden/bouncingball.sys tests/golden/bouncingball_chain_acc5.txt chain acc=5 domain=-10.1,10.1,-0.1,2.1 init=0.0,0.001,1.999,2.0 | tail -1
$T build/hostchain tests/golden/bouncingball.sys tests/golden/bouncingball_chain_acc6.txt chain acc=6 domain=-10.1,10.1,-0.1,2.1 init=-2.0,0.5,0.5,1.5 | tail -1
$T build/hostchain tests/golden/watertank.sys tests/golden/watertank_reach_acc3.txt upper acc=3 T=0.5 domain=0.0,1.0,4.5,9.0 loc=2 init=1.0,1.0,7.6,7.7 loc=1 init=0.0,0.0,5.7,5.8 | tail -1
$T build/hostchain2 tests/golden/springs.sys build/sp2.txt upper acc=2 T=40 domain=0.0,40.0,-2.0,2.0,-3.0,3.0,-1.0,3.0,0.0,4.0 init=0,0,0,0,0,0,0,0,3,3 | tail -1
$T build/hostchain2 tests/golden/boost.sys tests/golden/boost_reach_acc4.txt upper acc=4 T=0.011 domain=0.0,0.001,0.0,3.0,0.0,6.0 loc=1 init=0.0,0.0,1.0,1.0,1.0,1.0 | tail -1
