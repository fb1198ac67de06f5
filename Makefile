# Builds libnefes_b200.so (the C-ABI engine) for sm_100a.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $(EXTRA)
SRC := $(wildcard nefes_b200/csrc/*.cu)
OBJ := $(patsubst nefes_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := nefes_b200/lib/libnefes_b200.so

all: $(LIB)

build/%.o: nefes_b200/csrc/%.cu $(wildcard nefes_b200/csrc/*.cuh) include/nefes_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p nefes_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

clean:
	rm -rf build $(LIB)
.PHONY: all clean
