#!/bin/bash
# Developer tool: builds the single-thread emulation of the fourwf kernels (see emu_shim.h). Not a product path.
set -e
cd "$(dirname "$0")"
SRC=../../abinit_b200/csrc
g++ -O2 -std=c++17 -DABI_EMU -I. -I$SRC -x c++ -shared -fPIC -o libabinit_b200_emu.so \
   $SRC/fourwf.cu $SRC/plane_stage.cu $SRC/plane_inst_0.cu $SRC/plane_inst_1.cu $SRC/plane_inst_2.cu $SRC/plane_inst_3.cu $SRC/half_stage.cu $SRC/half_inst_0.cu $SRC/half_inst_1.cu $SRC/half_inst_2.cu $SRC/x_stage.cu $SRC/x_inst_0.cu $SRC/x_inst_1.cu $SRC/context.cu $SRC/api_fourwf.cu nonlop_stub.cpp -Wno-unused-function 2>&1 | grep -v "warning: ignoring" | head -50
