"""Host-side ingestion: pandas / pyarrow / dict sources -> HostFrame columns, categorical codes and
null handling, without touching the GPU."""
import numpy as np
import pytest

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
