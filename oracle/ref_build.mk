# Test infrastructure only (see oracle/README.md).
# Compiles the UNMODIFIED reference CPU sources where they lie under $(REF)
# (default /root/reference) with g++ directly -- the reference's own CMake
# build is not run -- and links them with oracle/ref_probe.cpp into
# oracle/_ref/gomc_probe_<ensemble>.  Flags follow the reference's GNU release
# flags (-O3 -m64 -fopenmp, CMakeLists.txt:34) and per-ensemble -DENSEMBLE=n
# (CMake/GOMCCPUSetup.cmake:4-17).  Nothing is copied out of $(REF); all
# outputs land in oracle/_ref/ (git-ignored, travels to the GPU box).
#
#   make -f oracle/ref_build.mk -j8            # NVT + GEMC probes
#   make -f oracle/ref_build.mk ENS=NVT -j8
REF      ?= /root/reference
OUT      ?= oracle/_ref
CXX      := /usr/bin/g++
CXXFLAGS := -O3 -m64 -fopenmp -std=c++17 -w
ENS_LIST ?= NVT GEMC NPT
ENSNUM_NVT  := 1
ENSNUM_GEMC := 2
ENSNUM_GCMC := 3
ENSNUM_NPT  := 4

INC := -I$(OUT)/include -I$(REF)/lib -I$(REF)/src -I$(REF)/src/cbmc -I$(REF)/src/moves \
       -I$(REF)/src/GPU

SRC := $(filter-out $(REF)/src/Main.cpp,$(wildcard $(REF)/src/*.cpp)) \
       $(wildcard $(REF)/src/cbmc/*.cpp) $(wildcard $(REF)/lib/*.cpp)

all: $(foreach e,$(ENS_LIST),$(OUT)/gomc_probe_$(e))

$(OUT)/include/GOMC_Config.h:
	@mkdir -p $(OUT)/include
	@printf '%s\n' '#define GOMC_VERSION_MAJOR 2' '#define GOMC_VERSION_MINOR 80' \
	  '#define GOMC_GTEST 0' '#define GOMC_GTEST_MPI 0' '#define GOMC_LIB_MPI 0' \
	  '#define GOMC_THREAD_MPI 0' '#define GOMC_MPI (GOMC_LIB_MPI || GOMC_THREAD_MPI)' \
	  '#define MPI_IN_PLACE_EXISTS 0' > $@

define ENS_RULES
OBJ_$(1) := $$(patsubst $(REF)/%.cpp,$(OUT)/obj_$(1)/%.o,$(SRC))
$(OUT)/obj_$(1)/%.o: $(REF)/%.cpp $(OUT)/include/GOMC_Config.h
	@mkdir -p $$(dir $$@)
	$(CXX) $(CXXFLAGS) -DENSEMBLE=$(ENSNUM_$(1)) $(INC) -c $$< -o $$@
$(OUT)/obj_$(1)/ref_probe.o: oracle/ref_probe.cpp $(OUT)/include/GOMC_Config.h
	@mkdir -p $$(dir $$@)
	$(CXX) $(CXXFLAGS) -DENSEMBLE=$(ENSNUM_$(1)) $(INC) -c $$< -o $$@
# (VolumeTransfer.h defines PrintAcceptKind out of line: the probe's copy is identical)
$(OUT)/gomc_probe_$(1): $$(OBJ_$(1)) $(OUT)/obj_$(1)/ref_probe.o
	$(CXX) $(CXXFLAGS) -Wl,--allow-multiple-definition $$^ -o $$@
endef
$(foreach e,NVT GEMC GCMC NPT,$(eval $(call ENS_RULES,$(e))))

# ---- the reference's shipped GPU build (performance comparator only) --------
# Same unmodified sources with -DGOMC_CUDA (CMake/GOMCCUDASetup.cmake:2-3) plus
# src/GPU/*.cu through nvcc with separable compilation, for sm_100.  It is NOT a
# parity oracle (its device arithmetic is not bit-equal to the CPU path); bench.py
# times it on the GPU box next to this repo's engine.
#   make -f oracle/ref_build.mk gpu -j8
CUDA_HOME ?= /usr/local/cuda
NVCC      := $(CUDA_HOME)/bin/nvcc
GPUARCH   ?= -gencode arch=compute_100,code=sm_100
GPUDEFS   := -DGOMC_CUDA -DENSEMBLE=1
CUSRC     := $(wildcard $(REF)/src/GPU/*.cu)
GOBJ      := $(patsubst $(REF)/%.cpp,$(OUT)/obj_GPU_NVT/%.o,$(SRC))
GCUOBJ    := $(patsubst $(REF)/%.cu,$(OUT)/obj_GPU_NVT/%.cu.o,$(CUSRC))
$(OUT)/obj_GPU_NVT/%.o: $(REF)/%.cpp $(OUT)/include/GOMC_Config.h
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) $(GPUDEFS) $(INC) -I$(CUDA_HOME)/include -c $< -o $@
$(OUT)/obj_GPU_NVT/%.cu.o: $(REF)/%.cu $(OUT)/include/GOMC_Config.h
	@mkdir -p $(dir $@)
	$(NVCC) -ccbin $(CXX) -O3 -std=c++17 -w -rdc=true $(GPUARCH) $(GPUDEFS) \
	  -Xcompiler -fopenmp $(INC) -c $< -o $@
$(OUT)/obj_GPU_NVT/ref_probe.o: oracle/ref_probe.cpp $(OUT)/include/GOMC_Config.h
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) $(GPUDEFS) $(INC) -I$(CUDA_HOME)/include -c $< -o $@
$(OUT)/gomc_probe_GPU_NVT: $(GOBJ) $(GCUOBJ) $(OUT)/obj_GPU_NVT/ref_probe.o
	$(NVCC) -ccbin $(CXX) $(GPUARCH) -Xcompiler -fopenmp $^ -o $@
gpu: $(OUT)/gomc_probe_GPU_NVT

clean:
	rm -rf $(OUT)
.PHONY: all gpu clean
