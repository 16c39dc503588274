#!/usr/bin/env bash
# Builds tests/hostsim/_build/fuzz: the drop-in host layer (host/*.hpp, unchanged) + fuzz.cpp, linked against the CPU stand-in
# of the C-ABI (dfsa_hostsim.cpp) INSTEAD of libdfsa_b200.so. TEST INFRASTRUCTURE ONLY -- see the header of dfsa_hostsim.cpp.
# The binary is git-ignored; it is a checker of host-side logic and never part of the product.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$(cd "$here/../.." && pwd)"
pkg="$root/distributed-full-state-algorithms_b200"
out="$here/_build"
mkdir -p "$out"
unset CC CXX
if [ ! -x "$out/fuzz" ] || [ -n "$(find "$pkg/host" "$root/include" "$here/fuzz.cpp" "$here/dfsa_hostsim.cpp" "$here/build.sh" -newer "$out/fuzz" -type f 2>/dev/null | head -1)" ]; then
    g++ -std=c++17 -O2 -Wall -I"$pkg/host" -I"$root/include" "$here/fuzz.cpp" "$here/dfsa_hostsim.cpp" -o "$out/fuzz"
    echo "hostsim: built $out/fuzz"
fi
# the extern "C" face of the host layer (host/dfsa_host_capi.cpp, what ctypes users and bench.py call) on the stand-in: lets a CPU
# test drive the per-state entry points that need a live StateVector (tests/hostsim/capi_on_standin.py)
if [ ! -f "$out/libdfsa_host_on_standin.so" ] || [ -n "$(find "$pkg/host" "$root/include" "$here/dfsa_hostsim.cpp" "$here/build.sh" -newer "$out/libdfsa_host_on_standin.so" -type f 2>/dev/null | head -1)" ]; then
    g++ -std=c++17 -O2 -Wall -fPIC -shared -I"$pkg/host" -I"$root/include" "$pkg/host/dfsa_host_capi.cpp" "$here/dfsa_hostsim.cpp" -o "$out/libdfsa_host_on_standin.so"
    echo "hostsim: built $out/libdfsa_host_on_standin.so"
fi
# the reference's own Catch2 suite (the objects tests/catch_dropin/build.sh compiled against host/*.hpp) linked against the stand-in
# instead of libdfsa_b200.so: its 20 cases then exercise the host layer's dispatch of every API function at up to 16 ranks on the CPU
catch="$here/../catch_dropin/_build"
if [ -f "$catch/tests.o" ] && [ -f "$catch/catch_amalgamated.o" ]; then
    if [ ! -x "$out/catch_on_standin" ] || [ "$catch/tests.o" -nt "$out/catch_on_standin" ] || [ "$here/dfsa_hostsim.cpp" -nt "$out/catch_on_standin" ]; then
        g++ -std=c++17 -O2 -I"$root/include" "$catch/tests.o" "$catch/catch_amalgamated.o" "$here/dfsa_hostsim.cpp" -o "$out/catch_on_standin"
        echo "hostsim: built $out/catch_on_standin"
    fi
fi
