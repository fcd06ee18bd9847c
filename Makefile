# In-tree build of the product library (sm_100a only) and nothing else.
NVCC ?= nvcc
CSRC := xped_b200/csrc
SRC := $(CSRC)/core.cpp $(CSRC)/layout.cpp $(CSRC)/ops.cpp $(CSRC)/extras.cpp $(CSRC)/ctm.cpp $(CSRC)/ad.cpp $(CSRC)/io.cpp $(CSRC)/dist_ops.cpp $(CSRC)/capi.cpp \
       $(CSRC)/kernels.cu $(CSRC)/gemm.cu $(CSRC)/linalg.cu $(CSRC)/precond.cu $(CSRC)/rsvd.cu $(CSRC)/dist.cu
HDR := $(wildcard $(CSRC)/*.hpp) include/xped_b200.h
LIB := xped_b200/lib/libxped_b200.so
OBJDIR := build/obj
OBJ := $(patsubst $(CSRC)/%,$(OBJDIR)/%.o,$(SRC))
FLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function

$(LIB): $(OBJ)
	mkdir -p xped_b200/lib
	$(NVCC) $(FLAGS) --shared -o $@ $(OBJ)

$(OBJDIR)/%.o: $(CSRC)/% $(HDR)
	mkdir -p $(OBJDIR)
	$(NVCC) $(FLAGS) -c -o $@ $<

clean:
	rm -rf $(LIB) $(OBJDIR)
