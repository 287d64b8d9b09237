# Build of libdiscregrid_b200.so (sm_100a only) and of the test-only oracle.
#   make lib     : discregrid_b200/lib/libdiscregrid_b200.so   (the product: CUDA kernels + C-ABI + host BVH builder)
#   make oracle  : oracle/liboracle.so (+ oracle/_ref when /root/reference is present)  -- test infrastructure
# -fmad=false / -ffp-contract=off are part of the numerical contract (bit-exact parity with the reference's
# non-FMA x86-64 build); dg_selftest() refuses to run a library built without them.
NVCC     ?= /usr/local/cuda/bin/nvcc
HOSTCXX  := $(shell command -v /usr/bin/g++ || echo g++)
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -ccbin $(HOSTCXX) -std=c++17 -O3 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden
CXXFLAGS := -std=c++17 -O3 -fPIC -ffp-contract=off -fvisibility=hidden
SRC      := discregrid_b200/csrc
OBJ      := build/obj
LIB      := discregrid_b200/lib/libdiscregrid_b200.so
CU       := k1_sdf k2_interp k3_density dg_api
HDRS     := $(wildcard $(SRC)/*.h $(SRC)/*.cuh) include/discregrid_b200.h

all: lib oracle
lib: $(LIB)

$(OBJ)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(OBJ)/bvh_build.o: $(SRC)/bvh_build.cpp $(HDRS)
	@mkdir -p $(OBJ)
	$(HOSTCXX) $(CXXFLAGS) -c $< -o $@

$(LIB): $(addprefix $(OBJ)/,$(addsuffix .o,$(CU))) $(OBJ)/bvh_build.o
	@mkdir -p $(dir $(LIB))
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -shared -o $@ $^ -Xlinker --exclude-libs=ALL

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean
.PHONY: all lib oracle clean
