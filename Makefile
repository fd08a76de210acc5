# Builds libeventflow.so (sm_100a only) and the C oracle helpers.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(EXTRA) -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v --expt-relaxed-constexpr
SRC       := $(wildcard event_flow_b200/csrc/*.cu)
OBJ       := $(patsubst event_flow_b200/csrc/%.cu,build/%.o,$(SRC))
LIB       := event_flow_b200/lib/libeventflow.so

all: $(LIB)

build/%.o: event_flow_b200/csrc/%.cu event_flow_b200/csrc/common.cuh event_flow_b200/csrc/tc_common.cuh include/eventflow.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIB): $(OBJ)
	@mkdir -p event_flow_b200/lib
	$(NVCC) -shared $(ARCH) -o $@ $(OBJ)

clean:
	rm -rf build event_flow_b200/lib

.PHONY: all clean
