# In-tree build of the product library (sm_100a only) and nothing else.
NVCC ?= nvcc
SRC := xped_b200/csrc/core.cpp xped_b200/csrc/layout.cpp xped_b200/csrc/ops.cpp xped_b200/csrc/extras.cpp \
       xped_b200/csrc/ctm.cpp xped_b200/csrc/capi.cpp xped_b200/csrc/kernels.cu xped_b200/csrc/gemm.cu xped_b200/csrc/linalg.cu xped_b200/csrc/precond.cu
HDR := xped_b200/csrc/core.hpp xped_b200/csrc/engine.hpp xped_b200/csrc/ctm.hpp xped_b200/csrc/precond.hpp include/xped_b200.h
LIB := xped_b200/lib/libxped_b200.so
FLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function --shared

$(LIB): $(SRC) $(HDR)
	mkdir -p xped_b200/lib
	$(NVCC) $(FLAGS) -o $@ $(SRC)

clean:
	rm -f $(LIB)
