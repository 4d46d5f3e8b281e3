"""The file catalog (npz / yaml catalog formats) against files WRITTEN BY THE REFERENCE's own
file_management package (tests/golden/catalog, oracle/make_golden.py: catalog_files): our loader reads them, and a
catalog written by our side has the reference's structure key for key."""
import os
import re
import shutil
import sys

import numpy as np
import pytest
import yaml

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "catalog")
CRIT = {'z_cut': 0.1, 'jet type': 'quark', 'fixed coupling': True, 'number of samples': 1000}
PRE = {'z_cut': 0.1, 'jet type': 'gluon', 'fixed coupling': False, 'number of samples': 1000}
SAMPLES = {'z_cut': 0.05, 'beta': 2, 'epsilon': 1e-10}


def _expected():
    rng = np.random.RandomState(11)
    edges = np.logspace(-6, 0, 21)
    crit = np.sort(rng.rand(21))[::-1].copy()
    xs, ys = np.logspace(-8, -1, 9), np.logspace(-8, 0, 7)
    pre = rng.rand(7, 7)
    return edges, crit, xs[:7], ys, pre, rng.rand(5)


@pytest.fixture()
def reference_catalog(tmp_path, monkeypatch):
    """a scratch copy of the reference-written catalog, with the working directory at its root (the stored
    filenames are relative to it, as the reference writes them)"""
    shutil.copytree(os.path.join(GOLDEN, "output"), tmp_path / "output")
    monkeypatch.chdir(tmp_path)
    from jetmontecarlo_b200.file_management import catalog_utils as CU
    CU.set_output_root("output", create=True)
    yield tmp_path
    CU.set_output_root("output", create=False)


def test_reads_what_the_reference_wrote(reference_catalog):
    from jetmontecarlo_b200.file_management.data_management import load_data, load_and_interpolate
    edges, crit, xs, ys, pre, samples = _expected()
    d = load_data('numerical integral', 'critical radiator', CRIT)
    np.testing.assert_array_equal(d['bins'], edges)
    np.testing.assert_array_equal(d['integral'], crit)
    d2 = load_data('numerical integral', 'pre-critical radiator', PRE)
    np.testing.assert_array_equal(d2['bins'], np.array([xs, ys]))
    np.testing.assert_array_equal(d2['integral'], pre)
    s = load_data('montecarlo samples', 'critical sudakov inverse transform', SAMPLES)
    np.testing.assert_array_equal(s['samples'], samples)
    # the interpolants of the stored integrals pass through the stored points
    f1 = load_and_interpolate('critical radiator', CRIT, monotonic=False, bounds=(edges[0], edges[-1]))
    np.testing.assert_allclose(f1(edges[3:-3]), crit[3:-3], rtol=1e-12)
    assert callable(load_and_interpolate('critical radiator', CRIT))            # monotone-domain form, the default
    f2 = load_and_interpolate('pre-critical radiator', PRE)
    np.testing.assert_allclose(f2(np.array([[xs[2], ys[4]]])), pre[2, 4], rtol=1e-12)
    with pytest.raises(FileNotFoundError):
        load_data('numerical integral', 'critical radiator', dict(CRIT, z_cut=0.2))


def test_writes_the_reference_format(reference_catalog):
    from jetmontecarlo_b200.file_management import catalog_utils as CU
    from jetmontecarlo_b200.file_management.data_management import save_new_data, load_data
    golden = yaml.safe_load(open(os.path.join(GOLDEN, "output", "examples", "file_catalog.yaml")))
    os.remove("output/examples/file_catalog.yaml")
    open("output/examples/file_catalog.yaml", "w").close()
    edges, crit, xs, ys, pre, samples = _expected()
    f_crit = save_new_data({'bins': edges, 'integral': crit}, 'numerical integral', 'critical radiator', CRIT, '.npz')
    save_new_data({'bins': np.array([xs, ys]), 'integral': pre}, 'numerical integral', 'pre-critical radiator', PRE, '.npz')
    save_new_data({'samples': samples, 'weights': np.ones(5)}, 'montecarlo samples',
                  'critical sudakov inverse transform', SAMPLES, '.pkl')
    mine = yaml.safe_load(open("output/examples/file_catalog.yaml"))
    strip = lambda name: re.sub(r"_[0-9a-f]{8}-[0-9a-f-]{27}\.", "_<uuid>.", name)
    assert sorted(mine) == sorted(golden)
    for data_type in golden:
        assert sorted(mine[data_type]) == sorted(golden[data_type])
        for source in golden[data_type]:
            assert sorted(mine[data_type][source]) == sorted(golden[data_type][source])          # the yaml keys
            for key, entry in golden[data_type][source].items():
                got = dict(mine[data_type][source][key])
                assert strip(got.pop('filename')) == strip(dict(entry)['filename'])
                assert got == {k: v for k, v in entry.items() if k != 'filename'}
    # same parameters again: the entry is replaced and the old file is gone
    f_new = save_new_data({'bins': edges, 'integral': crit * 2}, 'numerical integral', 'critical radiator', CRIT, '.npz')
    assert not os.path.exists(f_crit) and os.path.exists(f_new)
    np.testing.assert_array_equal(load_data('numerical integral', 'critical radiator', CRIT)['integral'], crit * 2)
    with pytest.raises(ValueError):
        save_new_data([1, 2], 'numerical integral', 'critical radiator', CRIT, '.npz')
    with pytest.raises(ValueError):
        save_new_data({}, 'numerical integral', 'critical radiator', CRIT, '.txt')
    with pytest.warns(UserWarning):
        CU.check_data_source('numerical integral', 'no such function')
    assert CU.dict_to_yaml_key({'b': 1, 'a': True}) == 'a : True | b : 1'


def test_install_as_registers_file_management():
    import jetmontecarlo_b200
    saved = {k: v for k, v in sys.modules.items()
             if k.split('.')[0] in ("jetmontecarlo", "file_management")}
    for k in saved:
        del sys.modules[k]
    try:
        aliased = jetmontecarlo_b200.install_as("jetmontecarlo")
        assert "file_management.data_management" in aliased
        from file_management.data_management import save_new_data, load_data, load_and_interpolate     # noqa: F401
        from file_management.catalog_utils import new_cataloged_filename, filename_from_catalog, fig_folder   # noqa: F401
    finally:
        for k in [k for k in sys.modules if k.split('.')[0] in ("jetmontecarlo", "file_management")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_loaders_and_scattered_radiator_rows(tmp_path, monkeypatch):
    """file_management/load_data.py:114-195 with the parameter dictionaries as arguments: 1-D radiator -> monotone
    interpolant, binned 2-D radiator -> grid, the subsequent radiator's scattered rows (one row of bins per
    theta_crit) -> a table whose nodes carry the nearest stored value of their own row; pickled functions come back."""
    monkeypatch.chdir(tmp_path)
    from jetmontecarlo_b200.file_management import catalog_utils as CU
    from jetmontecarlo_b200.file_management.data_management import save_new_data, load_and_tabulate
    from jetmontecarlo_b200.file_management.load_data import load_radiators, load_splittingfns
    CU.set_output_root("output")
    rp = {'jet type': 'quark', 'epsilon': 1e-10}
    edges = np.logspace(-10, 0, 30)
    save_new_data({'bins': edges, 'integral': np.linspace(5, 0, 30)}, 'numerical integral', 'critical radiator',
                  dict(rp, z_cut=.1), '.npz')
    xs, ys = np.logspace(-9, -1, 12), np.logspace(-8, 0, 9)
    save_new_data({'bins': np.array([xs[:9], ys]), 'integral': np.outer(np.linspace(3, 0, 9), np.linspace(2, 1, 9))},
                  'numerical integral', 'pre-critical radiator', dict(rp, z_cut=.1), '.npz')
    thetas = np.array([1., 1e-3, .1, 1e-2])                     # rows stored out of order
    rows_x = np.array([np.logspace(-12, 0, 30) * t ** 2 for t in thetas])
    rows_z = np.array([np.linspace(5, 0, 30) * (1 + j) for j in range(4)])
    save_new_data({'xs': rows_x.flatten(), 'ys': np.repeat(thetas, 30), 'integral': rows_z.flatten()},
                  'numerical integral', 'subsequent radiator', dict(rp, beta=2.), '.npz')
    save_new_data(np.sqrt, 'serialized function', 'splitting function', dict(rp, z_cut=.1), '.pkl')
    rads = load_radiators(rp, [.1], [2.], as_tables=True)
    np.testing.assert_allclose(rads['critical'][.1](edges[5:-5]), np.linspace(5, 0, 30)[5:-5], rtol=1e-12)
    assert rads['pre-critical'][.1].zs.shape == (9, 9) and rads['pre-critical'][.1].method == "RectangularGrid"
    table = rads['subsequent'][2.]
    assert table.method == "Nearest" and np.array_equal(table.ys, np.sort(thetas)) and table.zs.shape == (120, 4)
    for j, t in enumerate(np.sort(thetas)):
        row = list(thetas).index(t)
        for k in (0, 17, 60, 119):
            nearest = np.abs(rows_x[row] - table.xs[k]).argmin()
            assert table.zs[k, j] == rows_z[row][nearest]
    host = load_radiators(rp, [.1], [2.])                       # the reference's form: host interpolants
    assert callable(host['subsequent'][2.]) and callable(host['pre-critical'][.1])
    assert load_splittingfns(rp, [.1])[.1](4.) == 2.
    with pytest.raises(ValueError):
        load_and_tabulate('critical radiator', dict(rp, z_cut=.1))
