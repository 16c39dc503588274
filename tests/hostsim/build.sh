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
if [ -x "$out/fuzz" ] && [ -z "$(find "$pkg/host" "$root/include" "$here/fuzz.cpp" "$here/dfsa_hostsim.cpp" "$here/build.sh" -newer "$out/fuzz" -type f 2>/dev/null | head -1)" ]; then
    exit 0
fi
unset CC CXX
g++ -std=c++17 -O2 -Wall -I"$pkg/host" -I"$root/include" "$here/fuzz.cpp" "$here/dfsa_hostsim.cpp" -o "$out/fuzz"
echo "hostsim: built $out/fuzz"
