"""Saving / loading catalogued data and turning stored numerical integrals into interpolants.
ref: file_management/data_management.py:43-141.  Beyond the reference: ``load_and_tabulate`` hands a stored 2-D
radiator to the device-side conditional sampler (numerics/sudakov_sampling.py) as a ``TableRadiator``."""
import pickle as _std_pickle

import numpy as np

from . import catalog_utils as _catalog
from ..utils.interpolation import get_1d_interpolation, get_2d_interpolation

try:                        # the reference pickles with dill (lambdas, closures); plain pickle reads plain pickles
    import dill as pickle
except ImportError:         # pragma: no cover
    pickle = _std_pickle


class BackCompatUnpickler(pickle.Unpickler):
    """Unpickler that also looks classes up in the modules they moved to.  ref :17-37"""
    moved_to = ('jetmontecarlo.numerics.splitting', 'jetmontecarlo.numerics.radiators.generation')

    def find_class(self, module, name):
        for candidate in (module,) + self.moved_to:
            try:
                return super().find_class(candidate, name)
            except ModuleNotFoundError:
                continue
        raise ModuleNotFoundError("Could not find a valid module for loading the pickled object. "
                                  "Tried %s (given) and %s." % (module, list(self.moved_to)))


def save_new_data(data, data_type, data_source, params, extension):
    """Store ``data`` in a new catalogued file: '.pkl' (any object), '.npz' (dict of arrays) or '.npy'.  ref :43-69"""
    if extension not in ('.pkl', '.npz', '.npy'):
        raise ValueError("Extension must be .pkl, .npz, or .npy, not %s." % extension)
    if extension == '.npz' and not isinstance(data, dict):
        raise ValueError("Data must be a dictionary to be saved as a .npz file.")
    filename = _catalog.new_cataloged_filename(data_type, data_source, params, extension)
    if extension == '.pkl':
        with open(filename, 'wb') as file:
            pickle.dump(data, file)
    elif extension == '.npz':
        np.savez(filename, **{k: _host(v) for k, v in data.items()})
    else:
        np.save(filename, _host(data))
    return filename


def _host(value):
    """device tensors (the GPU path's results) are stored as the NumPy arrays the reference stores"""
    return value.detach().cpu().numpy() if hasattr(value, 'detach') else value


def load_data(data_type, data_source, params):
    """ref :72-93"""
    filename = str(_catalog.filename_from_catalog(data_type, data_source, params))
    extension = "." + filename.split('.')[-1]
    if extension == '.pkl':
        with open(filename, 'rb') as file:
            return BackCompatUnpickler(file).load()
    if extension in ('.npz', '.npy'):
        return np.load(filename, allow_pickle=True, mmap_mode='c')
    raise ValueError("Extension must be .pkl, .npz, or .npy,  %s." % extension)


def _grid_of(data, data_source):
    """(dimensionality, axes, values) of a stored numerical integral.  ref :103-123"""
    bins = data.get('bins')
    integral = data['integral']
    if bins is None:
        if data_source not in ('subsequent radiator',):
            raise ValueError("No bins found for data source %s." % data_source)
        return 2, (data.get('xs'), data.get('ys')), integral
    bins = np.asarray(bins)
    if bins.ndim == 2:
        return 2, (bins[0], bins[1]), integral
    return bins.ndim, (bins,), integral


def load_and_interpolate(data_source, params, **interp_kwargs):
    """The stored numerical integral as an interpolating function: 1-D monotone by default, 2-D on its grid.
    ref :96-141"""
    data = load_data('numerical integral', data_source, params)
    dim, axes, integral = _grid_of(data, data_source)
    if dim == 1:
        if interp_kwargs.get('monotonic') is None:
            interp_kwargs['monotonic'] = True
        return get_1d_interpolation(axes[0], integral, **interp_kwargs)
    if dim == 2:
        return get_2d_interpolation(axes[0], axes[1], integral, **interp_kwargs)
    raise ValueError('Dimensionality of data must be 1 or 2.')


def load_and_tabulate(data_source, params, method="RectangularGrid", num_x=None):
    """The stored 2-D numerical radiator as a ``TableRadiator`` for the device's conditional sampler.
    Binned data (pre-critical radiator: ``bins`` = two axes) is the table itself.  The subsequent radiator is stored
    as scattered points -- one row of observable bins per theta_crit, every row with its own bins
    (numerics/radiators/generation.py:253-400) -- and interpolated by nearest neighbour; it is re-tabulated here:
    its own theta_crit values form one axis, ``num_x`` (default 4 x the row length) log-spaced observable values
    over the rows' common range the other, and every node takes the value of the nearest stored point of its OWN row
    (the row's cumulative integral is a step function of the observable on that row's bins)."""
    from ..numerics.sudakov_sampling import TableRadiator
    data = load_data('numerical integral', data_source, params)
    dim, axes, integral = _grid_of(data, data_source)
    if dim != 2:
        raise ValueError('A radiator table is two-dimensional.')
    xs, ys, zs = np.asarray(axes[0], dtype=float), np.asarray(axes[1], dtype=float), np.asarray(integral, dtype=float)
    if zs.ndim == 2 and xs.ndim == 1 and zs.shape == (len(xs), len(ys)):
        return TableRadiator(xs, ys, zs, method)
    # scattered rows: ys is constant along a row
    thetas, first = np.unique(ys, return_index=True)
    row_len = len(ys) // len(thetas)
    if row_len * len(thetas) != len(ys) or not np.all(np.diff(np.sort(first)) == row_len):
        raise ValueError('scattered radiator data must be whole rows of constant theta')
    by_theta = np.argsort(ys.reshape(len(thetas), row_len)[:, 0])          # storage order -> increasing theta
    rows_x = xs.reshape(len(thetas), row_len)[by_theta]
    rows_z = zs.reshape(len(thetas), row_len)[by_theta]
    positive = rows_x[rows_x > 0]
    grid_x = np.logspace(np.log10(positive.min()), np.log10(rows_x.max()), int(num_x or 4 * row_len))
    table = np.empty((len(grid_x), len(thetas)))
    for j in range(len(thetas)):
        x_row = rows_x[j]
        right = np.clip(np.searchsorted(x_row, grid_x), 1, row_len - 1)
        nearest = np.where(np.abs(grid_x - x_row[right - 1]) <= np.abs(x_row[right] - grid_x), right - 1, right)
        table[:, j] = rows_z[j][nearest]
    return TableRadiator(grid_x, thetas, table, "Nearest")
