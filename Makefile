# Builds the C-ABI shared library of the hot path (sm_100a only) and the C oracle.
PKG      := road-object-detection-for-bdd100k_b200
CSRC     := $(PKG)/csrc
NVCC     ?= nvcc
NVFLAGS  := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
            -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v
SOURCES  := $(wildcard $(CSRC)/*.cu)
OBJECTS  := $(patsubst $(CSRC)/%.cu,build/%.o,$(SOURCES))
LIB      := $(PKG)/librodet_b200.so

all: $(LIB)

build/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/rodet_b200.h include/rodet_dlpack.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)
	@grep -E "error|warning" build/$*.ptxas.log || true

$(LIB): $(OBJECTS)
	$(NVCC) -shared -o $@ $(OBJECTS) -gencode arch=compute_100a,code=sm_100a -cudart shared

clean:
	rm -rf build $(LIB)

.PHONY: all clean
