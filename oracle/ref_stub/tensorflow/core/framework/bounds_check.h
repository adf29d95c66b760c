// empty stand-in: the reference's kernels only need TTypes<> (see framework/tensor.h)
