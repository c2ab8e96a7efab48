# Build libfgb200.so (CUDA kernels + C ABI + host-side LSSolver mirror) for sm_100a, in tree.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Iinclude
SRC := fibergen_b200/csrc
OBJ := build/obj
CU := $(wildcard $(SRC)/*.cu)
CPP := $(wildcard $(SRC)/*.cpp)
OBJS := $(patsubst $(SRC)/%.cu,$(OBJ)/%.o,$(CU)) $(patsubst $(SRC)/%.cpp,$(OBJ)/%.cpp.o,$(CPP))
LIB := fibergen_b200/libfgb200.so

all: $(LIB)

# geometric predicates of the phase initialisation: same products and sums as the reference's scalar code (no FMA contraction)
$(OBJ)/phase.o: NVFLAGS += -fmad=false

$(OBJ)/%.o: $(SRC)/%.cu $(wildcard $(SRC)/*.h) $(wildcard $(SRC)/*.cuh) include/fgb200.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJ)/%.cpp.o: $(SRC)/%.cpp $(wildcard $(SRC)/*.h) $(wildcard include/*.h)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart -ldl $(LDLIBS)

clean:
	rm -rf build $(LIB)
.PHONY: all clean
