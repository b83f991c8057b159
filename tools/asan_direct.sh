#!/bin/bash
# AddressSanitizer + UBSan over the host DIRECT driver (ibo_b200/csrc/direct.cpp) on the CPU: 60 random boxes (1..24 dims, fixed
# dims, corner optima, plateaus, maxsample cuts) in plain, sequential and speculating mode, which must agree with each other.
set -e
cd "$(dirname "$0")"
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -I../include -o /tmp/ibo_asan_direct asan_direct_main.cpp ../ibo_b200/csrc/direct.cpp
/tmp/ibo_asan_direct
