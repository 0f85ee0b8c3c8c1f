"""The few process-wide knobs of datashader_b200 (the reference has no flag system either)."""

# Return aggregates as CUDA tensors inside the DataArray instead of copying them to the host, the way the
# reference returns cupy-backed DataArrays for cudf sources (compiler.py:296-301).
device_results = False

# Record CUDA events around the fused aggregation launches (bench.py's roofline leg reads them).
time_kernels = False
kernel_events = []     # [(start_event, end_event)] appended per Canvas call when time_kernels is set

# K2: count() with the canvas privatised in shared memory (csrc/points.cu).  Used for resident chunks of at least
# `priv_min_rows` float32 or float64 points when the canvas fits (<= 954 000 cells: 226 KB of 2-bit fields, 97 % private); "off" forces the global-RED kernel.
priv_count = True
priv_min_rows = 1 << 24        # measured crossover vs global REDs: ~15-20 M rows (tools/bench_crossover.py)

# count() / by(cat, count()) on a u32 canvas of 1x..2x the L2 budget: one pass into 16-bit packed counters
# (dsb_points_count16) instead of two L2-banded passes.  l2_budget_bytes mirrors the library's default band budget.
count16 = True
count16_min_rows = 1 << 22
l2_budget_bytes = 96 << 20

# summary()-style plans (>= 3 accumulators) on small canvases: one specialised launch per accumulator group instead of
# one interpreted pass (pipeline._specialised_groups).
split_summary = True

# Bresenham lines with any / count / sum / max / min: the line kernel's own appends (the line's value held in a register)
# instead of the accumulator-plan interpreter, which re-reads the value column for every pixel.
lines_simple_path = True

# Single-accumulator plans (max / min / first / last of a float32 column, count) on canvases whose accumulator outgrows the
# L2 budget: bin the points into shared-memory-sized buckets and accumulate bucket by bucket (dsb_points_routed) instead of
# L2-banded passes with a global RED per hit.  routed_max_scratch_bytes bounds the record buffer (~8.8 bytes per point).
routed = True
routed_min_rows = 1 << 24
routed_max_scratch_bytes = 64 << 30
# first / last on canvases that DO fit L2: taken by dsb_points_routed as well once the frame holds this many rows per canvas cell
# (the library routes 10 rows per cell and only filters the rest; below 1.25 x that it would route every row)
routed_rows_per_cell_for_first = 16

# where(max | min) of a float32 selector on canvases beyond L2: the plain extreme first (routed), then a filtered pass that
# finds the row holding it (dsb_points_match32), instead of packed {key, row} atomics into an L2-banded 8-byte canvas.
where_two_pass = True

# max / min of a float32 column on an L2-resident canvas once the frame holds this many rows per canvas cell: the first
# `minmax_head_rows_per_cell` rows per cell go through dsb_points, the rest through dsb_points_minmax_rest (block thresholds in shared
# memory drop the rows that cannot win any more).  0 = off.
minmax_split_rows_per_cell = 512
minmax_head_rows_per_cell = 128
