"""Host-side ingestion: pandas / pyarrow / dict sources -> HostFrame columns, categorical codes and
null handling, without touching the GPU."""
import numpy as np
import pytest
import torch

from datashader_b200.frame import HostFrame, as_frame

pa = pytest.importorskip("pyarrow")


def test_arrow_table_columns_and_nulls():
    x = pa.array([0.1, 0.2, None, 0.4], type=pa.float32())
    i = pa.array([1, None, 3, 4], type=pa.int32())
    cat = pa.array(["a", "b", None, "a"]).dictionary_encode()
    t = pa.table({"x": x, "i": i, "cat": cat, "unused": pa.array(["p", "q", "r", "s"])})
    f = HostFrame.from_arrow(t, columns=["x", "i", "cat"], device="cpu")
    assert len(f) == 4 and set(f.columns) == {"x", "i", "cat"}
    np.testing.assert_array_equal(np.isnan(f.columns["x"].numpy()), [False, False, True, False])
    assert f.np_dtype("x") == np.float32
    assert f.np_dtype("i") == np.float64 and np.isnan(f.columns["i"].numpy()[1])      # pandas' promotion for int nulls
    assert f.categories["cat"] == ["a", "b"]
    np.testing.assert_array_equal(f.columns["cat"].numpy(), [0, 1, -1, 0])           # pandas' code for a missing category
    assert f.schema()["cat"] == ("categorical", ["a", "b"]) and f.schema()["x"] == ("float", None)
    with pytest.raises(ValueError, match="specified column not found"):
        HostFrame.from_arrow(t, columns=["nope"], device="cpu")
    with pytest.raises(ValueError, match="numeric or dictionary"):
        HostFrame.from_arrow(t, columns=["unused"], device="cpu")


def test_arrow_matches_pandas_ingestion():
    import pandas as pd
    rng = np.random.default_rng(0)
    df = pd.DataFrame({"x": rng.random(100, dtype=np.float32), "v": rng.normal(size=100),
                       "cat": pd.Categorical.from_codes(rng.integers(0, 3, 100), categories=["r", "g", "b"])})
    fp = as_frame(df, ["x", "v", "cat"], device="cpu")
    fa = as_frame(pa.Table.from_pandas(df), ["x", "v", "cat"], device="cpu")
    for c in ("x", "v", "cat"):
        np.testing.assert_array_equal(fp.columns[c].numpy(), fa.columns[c].numpy())
        assert fp.np_dtype(c) == fa.np_dtype(c)
    assert fp.categories == fa.categories


def test_chunked_arrow_and_parquet_roundtrip(tmp_path):
    import pyarrow.parquet as pq
    a = pa.chunked_array([pa.array([1.0, 2.0]), pa.array([3.0])])
    t = pa.table({"x": a, "y": pa.array([4.0, 5.0, 6.0])})
    f = HostFrame.from_arrow(t, device="cpu")
    np.testing.assert_array_equal(f.columns["x"].numpy(), [1.0, 2.0, 3.0])
    pq.write_table(t, tmp_path / "p.parquet")
    g = HostFrame.from_parquet(tmp_path / "p.parquet", columns=["y"], device="cpu")
    assert set(g.columns) == {"y"} and len(g) == 3


def test_dict_source():
    f = as_frame({"x": np.arange(4, dtype=np.float32), "y": np.arange(4.0), "z": np.zeros(4)}, ["x", "y"], device="cpu")
    assert isinstance(f, HostFrame) and set(f.columns) == {"x", "y"} and f.np_dtype("x") == np.float32
    with pytest.raises(ValueError, match="specified column not found"):
        as_frame({"x": np.arange(4.0)}, ["x", "y"], device="cpu")
    with pytest.raises(ValueError, match="source must be a pandas or dask DataFrame"):
        as_frame([1, 2, 3], ["x"], device="cpu")


def test_ragged_array_container_and_frame():
    """RaggedArray (the input type of the ragged line / area glyphs, reference datatypes.py:209-300) as a pandas column, and its
    trip into a HostFrame: flat values + int64 start index per row, schema kind 'ragged'."""
    import pandas as pd
    from datashader_b200.datatypes import RaggedArray, RaggedDtype
    from datashader_b200.frame import HostFrame, RaggedColumn
    ra = RaggedArray([[0.0, 1, 2], [1.0, 3, 4, 2], None, [0.5]], dtype="float32")
    assert isinstance(ra.dtype, RaggedDtype) and ra.dtype.name == "ragged[float32]" and len(ra) == 4
    np.testing.assert_array_equal(ra.start_indices, [0, 3, 7, 7])
    np.testing.assert_array_equal(ra.flat_array, np.array([0, 1, 2, 1, 3, 4, 2, 0.5], np.float32))
    np.testing.assert_array_equal(ra.isna(), [False, False, True, False])
    np.testing.assert_array_equal(ra[1], np.array([1, 3, 4, 2], np.float32))
    df = pd.DataFrame({"x": ra, "v": [1.0, 2.0, 3.0, 4.0]})
    sub = df.iloc[1:4]["x"].array                      # slicing re-bases the start indices
    np.testing.assert_array_equal(sub.start_indices, [0, 4, 4])
    np.testing.assert_array_equal(pd.concat([df, df])["x"].array.start_indices, [0, 3, 7, 7, 8, 11, 15, 15])
    assert pd.Series([[1, 2], [3]], dtype="ragged[float64]").array.flat_array.dtype == np.float64
    wrapped = RaggedArray({"start_indices": np.array([0, 2]), "flat_array": np.arange(5.0)})
    np.testing.assert_array_equal(wrapped[1], [2.0, 3.0, 4.0])
    with pytest.raises(ValueError):
        RaggedArray({"start_indices": np.array([3, 1]), "flat_array": np.arange(5.0)})
    hf = HostFrame.from_pandas(df, columns=["x", "v"])
    assert hf.schema()["x"] == ("ragged", None) and hf.schema()["v"][0] == "float" and len(hf) == 4 and hf.n_chunks() == 1
    col = hf.columns["x"]
    assert isinstance(col, RaggedColumn) and col.starts.dtype == torch.int64 and col.flat.dtype == torch.float32
    np.testing.assert_array_equal(col.starts.numpy(), [0, 3, 7, 7])

    class Foreign:                                     # the reference's own RaggedArray is read by attribute
        flat_array = np.arange(4.0)
        start_indices = np.array([0, 1], dtype=np.uint8)
    hf2 = HostFrame({"x": Foreign(), "v": np.zeros(2)})
    assert hf2.schema()["x"][0] == "ragged" and hf2.columns["x"].starts.tolist() == [0, 1]


def test_ragged_glyph_selection_and_validation():
    """Canvas.line / Canvas.area with axis=1 and scalar column names select the ragged glyphs (core.py:443-444, 676-677, 698-699)
    and reject non-ragged columns with the reference's messages (line.py:458-467, area.py:917-926) before any kernel runs."""
    import pandas as pd
    import datashader_b200 as ds
    from datashader_b200.datatypes import RaggedArray
    df = pd.DataFrame({"x": RaggedArray([[0.0, 1.0], [0.5]]), "y": [1.0, 2.0]})
    cvs = ds.Canvas(plot_width=8, plot_height=8, x_range=(0, 1), y_range=(0, 1))
    with pytest.raises(ValueError, match="y must be a RaggedArray"):
        cvs.line(df, "x", "y", axis=1)
    with pytest.raises(ValueError, match="x must be a RaggedArray"):
        cvs.area(df, "y", "x", axis=1)
    with pytest.raises(ValueError, match="y must be a RaggedArray"):
        cvs.area(df, "x", "x", axis=1, y_stack="y")
