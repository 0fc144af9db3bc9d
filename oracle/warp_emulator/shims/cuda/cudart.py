"""Inert stand-in for cuda-python's cudart (CUDA-GL interop in the reference constructor, object.py:100-113)."""


class cudaError_t:
    cudaSuccess = None


cudaError_t.cudaSuccess = cudaError_t()


class cudaGraphicsRegisterFlags:
    cudaGraphicsRegisterFlagsWriteDiscard = 0


def cudaGraphicsGLRegisterBuffer(*a): return (cudaError_t.cudaSuccess, 0)
def cudaGraphicsMapResources(*a): return (cudaError_t.cudaSuccess,)
def cudaGraphicsResourceGetMappedPointer(*a): return (cudaError_t.cudaSuccess, 0, 0)
def cudaGraphicsUnmapResources(*a): return (cudaError_t.cudaSuccess,)
