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
CU       := k1_sdf k2_interp k3_density k4_reduce dg_api
HDRS     := $(wildcard $(SRC)/*.h $(SRC)/*.cuh) include/discregrid_b200.h

all: lib cpp oracle
lib: $(LIB)

# C++ facade (cpp/include/Discregrid/*: the reference's class API over the C-ABI) -- the reference tools and a checker.
# Eigen3 is not installed in this image: cpp/eigen_min is used unless EIGEN3_INCLUDE_DIR points at the real one.
EIGEN3_INCLUDE_DIR ?= cpp/eigen_min
CPPBIN   := build/bin
CPPFLAGS := -std=c++11 -O2 -ffp-contract=off -I$(EIGEN3_INCLUDE_DIR) -Icpp/include -Iinclude
CPPLINK  := -Ldiscregrid_b200/lib -ldiscregrid_b200 -Wl,-rpath,'$$ORIGIN/../../discregrid_b200/lib'
CPPHDRS  := $(wildcard cpp/include/Discregrid/* cpp/include/Discregrid/*/*) include/discregrid_b200.h
cpp: $(CPPBIN)/GenerateSDF $(CPPBIN)/GenerateDensityMap $(CPPBIN)/DiscreteFieldToBitmap $(CPPBIN)/facade_check $(CPPBIN)/bvh_host_check $(CPPBIN)/reduce_facade_check $(CPPBIN)/sort_replay_check $(CPPBIN)/fast_div_check $(CPPBIN)/libk1emu.so $(CPPBIN)/libk1emu_knobs.so $(CPPBIN)/libk1emu_wave.so $(CPPBIN)/libk1emu_perlane.so $(CPPBIN)/libk1emu_packet.so $(CPPBIN)/libk23emu.so $(CPPBIN)/libk23emu_knobs.so $(CPPBIN)/libdgemu.so
$(CPPBIN)/DiscreteFieldToBitmap: cpp/cmd/discrete_field_to_bitmap.cpp $(CPPHDRS) $(LIB)
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) $(CPPFLAGS) $< -o $@ $(CPPLINK)
$(CPPBIN)/bvh_host_check: tests/cpp/bvh_host_check.cpp $(SRC)/bvh_build.cpp $(SRC)/host_threads.cpp $(SRC)/sort_replay.cpp $(SRC)/sort_replay.h $(SRC)/bvh_build.h
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) $(CXXFLAGS) -fvisibility=default tests/cpp/bvh_host_check.cpp $(SRC)/bvh_build.cpp $(SRC)/host_threads.cpp $(SRC)/sort_replay.cpp -o $@ -lpthread
# test libraries: the K1 translation unit (kernels + launchers) compiled for the CPU by tests/emu -- default knobs, and the prepared
# the knobs that are ON by default (redux vote, per-array brick shape) switched off, so that both settings stay checked
K1EMU_SRC := tests/emu/k1_emu.cpp $(SRC)/bvh_build.cpp $(SRC)/host_threads.cpp $(SRC)/sort_replay.cpp
K1EMU_DEP := $(K1EMU_SRC) tests/emu/cuda_emu.h $(SRC)/k1_sdf.cu $(HDRS)
CUDA_INC  ?= /usr/local/cuda/include
$(CPPBIN)/libk1emu.so: $(K1EMU_DEP)
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) -std=c++17 -O2 -ffp-contract=off -fPIC -shared -I$(CUDA_INC) -Itests/emu -I$(SRC) $(K1EMU_SRC) -o $@ -lpthread
$(CPPBIN)/libk1emu_knobs.so: $(K1EMU_DEP)
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) -std=c++17 -O2 -ffp-contract=off -fPIC -shared -DK1_PACKET=0 -DK1_VOTE_REDUX=0 -DK1_BRICK_AUTO=0 -I$(CUDA_INC) -Itests/emu -I$(SRC) $(K1EMU_SRC) -o $@ -lpthread
# the wavefront node-loop kernel (K1_WAVE=1) and the per-lane one (K1_WAVE=0), whichever is not the default build, stay checked
$(CPPBIN)/libk1emu_wave.so: $(K1EMU_DEP)
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) -std=c++17 -O2 -ffp-contract=off -fPIC -shared -DK1_WAVE=1 -I$(CUDA_INC) -Itests/emu -I$(SRC) $(K1EMU_SRC) -o $@ -lpthread
$(CPPBIN)/libk1emu_packet.so: $(K1EMU_DEP)
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) -std=c++17 -O2 -ffp-contract=off -fPIC -shared -DK1_PACKET=1 -I$(CUDA_INC) -Itests/emu -I$(SRC) $(K1EMU_SRC) -o $@ -lpthread
$(CPPBIN)/libk1emu_perlane.so: $(K1EMU_DEP)
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) -std=c++17 -O2 -ffp-contract=off -fPIC -shared -DK1_WAVE=0 -DK1_PACKET=0 -I$(CUDA_INC) -Itests/emu -I$(SRC) $(K1EMU_SRC) -o $@ -lpthread
K23EMU_SRC := tests/emu/k2_emu.cpp tests/emu/k3_emu.cpp
K23EMU_DEP := $(K23EMU_SRC) tests/emu/cuda_emu.h $(SRC)/k2_interp.cu $(SRC)/k3_density.cu $(HDRS)
$(CPPBIN)/libk23emu.so: $(K23EMU_DEP)
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) -std=c++17 -O2 -ffp-contract=off -fPIC -shared -I$(CUDA_INC) -Itests/emu -I$(SRC) $(K23EMU_SRC) -o $@ -lpthread
$(CPPBIN)/libk23emu_knobs.so: $(K23EMU_DEP)
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) -std=c++17 -O2 -ffp-contract=off -fPIC -shared -DK3_FAST_DIV=0 -I$(CUDA_INC) -Itests/emu -I$(SRC) $(K23EMU_SRC) -o $@ -lpthread
# test library: the WHOLE product library with every kernel emulated and the CUDA runtime stubbed on host memory -- lets the Python-level
# and tool-level `-m gpu` tests be rehearsed on the CPU (tests/test_gpu_rehearsal.py); never loaded by the product
DGEMU_SRC := tests/emu/dgapi_emu.cpp tests/emu/k1_emu.cpp tests/emu/k2_emu.cpp tests/emu/k3_emu.cpp tests/emu/k4_emu.cpp tests/emu/cudart_stub.cpp \
             $(SRC)/bvh_build.cpp $(SRC)/host_threads.cpp $(SRC)/reduce_field.cpp $(SRC)/obj_reader.cpp $(SRC)/sort_replay.cpp
$(CPPBIN)/libdgemu.so: $(DGEMU_SRC) tests/emu/cuda_emu.h $(wildcard $(SRC)/*.cu) $(HDRS)
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) -std=c++17 -O2 -ffp-contract=off -fPIC -shared -Wl,-Bsymbolic -Wno-subobject-linkage -I$(CUDA_INC) -Itests/emu -I$(SRC) -Iinclude $(DGEMU_SRC) -o $@ -lpthread
# test binary: the reciprocal-based exact division of fast_div.h against '/' (brute force)
$(CPPBIN)/fast_div_check: tests/cpp/fast_div_check.cpp $(SRC)/fast_div.h $(SRC)/dg_device.cuh
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) $(CXXFLAGS) -I$(SRC) tests/cpp/fast_div_check.cpp -o $@
# test binary: the threaded std::sort replay of reduce_field.cpp against std::sort (ties, depth exhaustion)
$(CPPBIN)/sort_replay_check: tests/cpp/sort_replay_check.cpp $(SRC)/reduce_field.cpp $(SRC)/host_threads.cpp $(SRC)/sort_replay.cpp $(SRC)/sort_replay.h $(SRC)/reduce_field.h $(SRC)/dg_device.cuh
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) $(CXXFLAGS) -fvisibility=default -I$(SRC) tests/cpp/sort_replay_check.cpp $(SRC)/reduce_field.cpp $(SRC)/host_threads.cpp $(SRC)/sort_replay.cpp -o $@ -lpthread
# test binary: the facade's reduceField with dg_node_positions interposed by the oracle (runs without a GPU)
$(CPPBIN)/reduce_facade_check: tests/cpp/reduce_facade_check.cpp $(CPPHDRS) $(LIB) oracle/liboracle.so
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) $(CPPFLAGS) $< -o $@ $(CPPLINK) -Loracle -loracle -Wl,-rpath,'$$ORIGIN/../../oracle'
oracle/liboracle.so:
	$(MAKE) -C oracle oracle
$(CPPBIN)/GenerateSDF: cpp/cmd/generate_sdf.cpp $(CPPHDRS) $(LIB)
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) $(CPPFLAGS) $< -o $@ $(CPPLINK)
$(CPPBIN)/GenerateDensityMap: cpp/cmd/generate_density_map.cpp $(CPPHDRS) $(LIB)
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) $(CPPFLAGS) $< -o $@ $(CPPLINK)
$(CPPBIN)/facade_check: cpp/cmd/facade_check.cpp $(CPPHDRS) $(LIB)
	@mkdir -p $(CPPBIN)
	$(HOSTCXX) $(CPPFLAGS) $< -o $@ $(CPPLINK)

$(OBJ)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(OBJ)/%.o: $(SRC)/%.cpp $(HDRS)
	@mkdir -p $(OBJ)
	$(HOSTCXX) $(CXXFLAGS) -c $< -o $@

$(LIB): $(addprefix $(OBJ)/,$(addsuffix .o,$(CU))) $(OBJ)/bvh_build.o $(OBJ)/host_threads.o $(OBJ)/reduce_field.o $(OBJ)/obj_reader.o $(OBJ)/sort_replay.o
	@mkdir -p $(dir $(LIB))
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -shared -o $@ $^ -Xlinker --exclude-libs=ALL

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean
.PHONY: all lib cpp oracle clean
