#!/bin/sh
# development aid: host builds of the kernel headers (tests/tools); not part of the product
set -e
cd "$(dirname "$0")/../.."
mkdir -p build
F="-std=c++17 -O1 -g -frounding-math -Wall -Wno-unused"
S="release-1.0_b200/csrc/ab200_system.cpp release-1.0_b200/csrc/ab200_paving.cpp"
g++ $F -o build/hostdbg tests/tools/hostdbg.cpp $S
g++ $F -DAB_KW=2 -o build/hostdbg2 tests/tools/hostdbg.cpp $S
g++ $F -o build/hoststep tests/tools/hoststep.cpp release-1.0_b200/csrc/ab200_system.cpp
g++ $F -o build/hostlap tests/tools/hostlap.cpp $S
g++ $F -x c++ -o build/hostchain tests/tools/hostchain.cpp $S
g++ $F -o build/hostkat tests/tools/hostkat.cpp
