# cuSten-B200 build: sm_100a only.
#   make lib      -> custen_b200/lib/libcuSten.a        (relocatable device code, the reference's library form; also
#                                                        carries the additive C entry points of api_c.cu / slab.cu)
#                    custen_b200/lib/libcusten_b200.so   (C ABI, device-linked, for FFI users and the tests)
#   make oracle   -> oracle/_ref/*                       (CPU oracle + the reference rebuilt from /root/reference)
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -rdc=true -Xcompiler -fPIC -Xcompiler -fvisibility=default
SRCDIR    := custen_b200/csrc
OBJDIR    := build/obj
LIBDIR    := custen_b200/lib

CORE_SRC  := kernels.cu plan.cu api_cpp.cu
CABI_SRC  := api_c.cu slab.cu cahn.cu cahn_part.cu pent_tma.cu pent_part.cu
CORE_OBJ  := $(patsubst %.cu,$(OBJDIR)/%.o,$(CORE_SRC))
CABI_OBJ  := $(patsubst %.cu,$(OBJDIR)/%.o,$(CABI_SRC))
HEADERS   := $(wildcard $(SRCDIR)/*.h $(SRCDIR)/*.cuh include/*.h)

.PHONY: all lib oracle examples clean
all: lib oracle examples

lib: $(LIBDIR)/libcuSten.a $(LIBDIR)/libcusten_b200.so

$(OBJDIR)/%.o: $(SRCDIR)/%.cu $(HEADERS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

# the archive = the reference's API + the additive C ABI and the multi-GPU slab layer (both only need the core), so that
# a C++ program with its own __device__ functions can device-link custen_mg_* too (examples/multi_gpu_stencil.cu)
ARCHIVE_OBJ := $(CORE_OBJ) $(OBJDIR)/api_c.o $(OBJDIR)/slab.o
$(LIBDIR)/libcuSten.a: $(ARCHIVE_OBJ)
	@mkdir -p $(LIBDIR)
	rm -f $@
	$(NVCC) --lib $(ARCHIVE_OBJ) --output-file $@

$(LIBDIR)/libcusten_b200.so: $(CORE_OBJ) $(CABI_OBJ)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(CORE_OBJ) $(CABI_OBJ)

# the re-hosted Cahn-Hilliard driver program (reference: cuPentSpeedUp/cuPentCahnADITiming)
examples: examples/bin/cuPentCahnADI examples/bin/registered_fun examples/bin/multi_gpu_stencil

# a user program against the C++ drop-in API and the static archive, with a registered __device__ function
examples/bin/registered_fun: examples/registered_fun.cu $(LIBDIR)/libcuSten.a $(HEADERS)
	@mkdir -p examples/bin
	$(NVCC) $(ARCH) -O3 -std=c++17 -rdc=true -o $@ $< $(LIBDIR)/libcuSten.a

examples/bin/multi_gpu_stencil: examples/multi_gpu_stencil.cu $(LIBDIR)/libcuSten.a $(HEADERS)
	@mkdir -p examples/bin
	$(NVCC) $(ARCH) -O3 -std=c++17 -rdc=true -o $@ $< $(LIBDIR)/libcuSten.a

examples/bin/cuPentCahnADI: examples/cuPentCahnADI.cu $(LIBDIR)/libcusten_b200.so
	@mkdir -p examples/bin
	$(NVCC) $(ARCH) -O3 -std=c++17 -o $@ $< -L$(LIBDIR) -lcusten_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../../$(LIBDIR)'

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIBDIR) oracle/_ref examples/bin
