"""Process-wide switches of the b200 sparseconvnet surface."""
import contextlib
import os

# B200SCN_MATH_* of include/b200scn.h: 0 = fp32 FFMA, 1 = tf32x3 tensor-core (fp32-equivalent), 2 = tf32
MATH_FP32, MATH_TF32X3, MATH_TF32 = 0, 1, 2
math_mode = MATH_FP32

# False: eager regime (exact row counts, one host read-back per level).
# True : static regime (capacity-sized matrices, device-side row counts, CUDA-graph capturable).
static_shapes = False
# True: weight-gradient kernels run on a side stream, overlapping the dgrad chain (forked and
# joined with events, so it is capturable); callers must call functional.join_side_streams()
# before consuming parameter gradients.
overlap_wgrad = False

# True: forward / dgrad of SubmanifoldConvolution, Convolution and NetworkInNetwork run on the tcgen05 tensor cores as
# a rule-table-driven implicit GEMM (csrc/sparse_tc.cu) whenever the channel counts are multiples of 16; False keeps
# the mma.sync pair-list kernels for them (they still serve Deconvolution and every weight gradient)
use_tc = not os.environ.get("B200SCN_NO_TC")
# narrowest layer (min(nIn, nOut)) the tcgen05 kernels take; below it the mma.sync pair-list kernels are faster
tc_min_channels = int(os.environ.get("B200SCN_TC_MIN_C", "96"))
tc_wgrad_min_channels = int(os.environ.get("B200SCN_TC_WGRAD_MIN_C", "96"))
# True: their weight gradients too (k_sparse_wgrad_tc); False: mma.sync over the forward pair lists (wgrad_gpl.cu)
use_tc_wgrad = not os.environ.get("B200SCN_NO_TC_WGRAD")

# True: ConcatTable(shortcut | ... SubmanifoldConvolution) + AddTable -- the residual blocks of scn.UNet -- form their
# sum inside the last convolution's epilogue instead of a separate elementwise pass
fuse_residual = not os.environ.get("B200SCN_NO_FUSE_RES")

# True: tensor-core convolutions read precomputed grouped pair lists (csrc/conv_gpl.cu) instead of
# re-listing the rule pairs inside every launch (csrc/conv_pairs.cu)
use_gpl = not os.environ.get("B200SCN_NO_GPL")
# levels with fewer rows keep the in-kernel listing (conv_pairs.cu, offset-split CTAs): building the lists
# costs more than it saves when one launch is a few microseconds
gpl_min_rows = int(os.environ.get("B200SCN_GPL_MIN_ROWS", "0"))

# rows per tile of the grouped pair lists (64 / 128); None = the library's recommendation for the level size
gpl_tile_rows = int(os.environ["B200SCN_GPL_TILE_ROWS"]) if os.environ.get("B200SCN_GPL_TILE_ROWS") else None

# True: weight gradients walk the forward pair list (csrc/wgrad_gpl.cu) instead of the neighbour table
wgrad_gpl = not os.environ.get("B200SCN_NO_WGRAD_GPL")

# True: convolutions over grouped pair lists read a fragment-ordered copy of the weights (functional.packed_weights)
pack_weights = not os.environ.get("B200SCN_NO_PACK")

# static regime: tile rows per spatial size.  A level's capacity says nothing about how many rows it really
# has, so the trainer calibrates this once on a real event (Trainer._calibrate_tiles)
gpl_tile_rows_by_size = {}
gpl_rows_by_size = {}  # the calibrated row counts themselves (launch heuristics)

# static regime: levels below this capacity build their pair lists in one pass into fixed-stride slots
gpl_fixed_stride_max_rows = 128 * 8 * 148

# static regime: the rulebooks and pair lists of the COARSER levels are built on a side stream right after the input
# level exists, while the main stream already runs the level-0 layers (metadata.Metadata._prefetch).  The list of
# builds is recorded during the first static forward of a given (dimension, input size) and replayed afterwards.
prefetch_rules = os.environ.get("B200SCN_PREFETCH_RULES", "1") != "0"
md_plans = {}

# set by a trainer: called with a tag when the backward pass of a MarkedSequential (the part of scn.UNet below
# the first level) has completed -- the point where most gradient bytes exist and their all-reduce can start
backward_mark_callback = None

# static regime only: int32 device tensor[1] holding the number of valid input rows (None = all rows)
static_input_rows = None


@contextlib.contextmanager
def static_mode(enabled=True):
    global static_shapes
    prev = static_shapes
    static_shapes = enabled
    try:
        yield
    finally:
        static_shapes = prev


@contextlib.contextmanager
def math(mode):
    global math_mode
    prev = math_mode
    math_mode = mode
    try:
        yield
    finally:
        math_mode = prev
