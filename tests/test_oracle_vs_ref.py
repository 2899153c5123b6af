"""CPU: the oracle against the reference's own host code compiled from its sources into
oracle/_ref/libref_host.so (only where that library exists -- it is built in the container that
mounts /root/reference and travels to the GPU box as a prebuilt file)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref(orc):
    if not orc.ref_host_available():
        pytest.skip("oracle/_ref/libref_host.so not built (reference tree absent)")
    return orc.RefHost()


@pytest.mark.parametrize("n", [2, 3, 6, 11, 40])
def test_writer_reader_builder_identical(orc, ref, n, tmp_path):
    pr, po = str(tmp_path / "ref.mtx"), str(tmp_path / "orc.mtx")
    ref.write_stencil(n, pr)
    orc.write_mtx_stencil5(n, po)
    assert open(pr, "rb").read() == open(po, "rb").read()
    m, ent_ref = ref.load(pr)
    rows, cols, nnz, grid, ent = orc.load_mtx(po)
    assert (m.rows, m.cols, m.nnz, m.grid_size) == (rows, cols, nnz, grid)
    assert ent_ref.tobytes() == ent.tobytes()
    rp_r, ci_r, va_r = ref.build_csr(m)
    rp, ci, va = orc.build_csr(rows, cols, ent)
    assert np.array_equal(rp_r, rp) and np.array_equal(ci_r, ci) and np.array_equal(va_r, va)


def test_builder_on_shuffled_duplicates(orc, ref, tmp_path):
    """non-stencil input: random order, duplicate (row, col) pairs -> stable per-row column sort."""
    rng = np.random.default_rng(42)
    rows, nnz = 37, 400
    r = rng.integers(0, rows, nnz)
    c = rng.integers(0, rows, nnz)
    v = rng.uniform(-1, 1, nnz)
    p = str(tmp_path / "rand.mtx")
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (rows, rows, nnz))
        for k in range(nnz):
            f.write("%d %d %.17g\n" % (r[k] + 1, c[k] + 1, v[k]))
    m, ent_ref = ref.load(p)
    rows2, cols2, nnz2, grid, ent = orc.load_mtx(p)
    assert grid == -1 and ent_ref.tobytes() == ent.tobytes()
    a = ref.build_csr(m)
    b = orc.build_csr(rows2, cols2, ent)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
