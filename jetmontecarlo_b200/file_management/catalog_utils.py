"""The YAML file catalog: {data type: {data source: {"k1 : v1 | k2 : v2 | ...": {params as strings, filename}}}}.
ref: file_management/catalog_utils.py:1-236 (same module-level names; the folders are module attributes and can be
re-pointed with ``set_output_root``)."""
import time
import uuid
import warnings
from pathlib import Path

import yaml

# ---- folders (ref :15-22): module attributes, as in the reference, but all derived from one root ----
def _layout(root):
    """(figures, output folder, storage, trash bin, catalog file) below ``root``"""
    root = Path(root)
    out = root / 'examples'
    return root / 'figures' / 'current', out, out / 'current', out / 'overwritten', out / 'file_catalog.yaml'


fig_folder, example_output_folder, example_storage, example_trashbin, example_catalog = _layout('output')


def set_output_root(root, create=True):
    """Point the catalog, its storage and its trash bin below ``root`` (the reference hard-codes ``output/`` relative
    to the working directory); creates the folders and an empty catalog file on request."""
    global fig_folder, example_output_folder, example_storage, example_trashbin, example_catalog
    fig_folder, example_output_folder, example_storage, example_trashbin, example_catalog = _layout(root)
    if create:
        for folder in (fig_folder, example_storage, example_trashbin):
            folder.mkdir(parents=True, exist_ok=True)
        example_catalog.touch()


# ---- the vocabulary of the catalog (ref :33-52) ----
recognized_data_types = ['montecarlo samples', 'numerical integral', 'serialized function']
emission_types = ['ungroomed', 'critical', 'subsequent', 'pre-critical']
function_types = ['radiator', 'sudakov function']
all_functions = ['%s %s' % (e, f) for e in emission_types for f in function_types] + ['splitting function']
recognized_data_sources = {
    'montecarlo samples': [e + ' phase space' for e in emission_types] + ['parton shower']
                          + [e + ' sudakov inverse transform' for e in emission_types],
    'numerical integral': all_functions,
    'serialized function': all_functions,
}


def _complain(ok, message, warn_only):
    if ok:
        return
    if warn_only:
        warnings.warn(message)
    else:
        raise AssertionError(message)


def check_data_type(data_type, warn_only=True):
    """ref :56-62"""
    _complain(data_type in recognized_data_types, "Unrecognized data type: %s" % data_type, warn_only)


def check_data_source(data_type, data_source, warn_only=True):
    """ref :65-75"""
    _complain(data_source in (recognized_data_sources.get(data_type) or ()),
              "Unrecognized data source %s from data type %s." % (data_source, data_type), warn_only)


def check_params(params, warn_only=True):
    """no ``None`` (or "None") among the parameter values.  ref :77-88"""
    values = list(params.values())
    _complain(not any(v is None or (isinstance(v, str) and v == "None") for v in values),
              "There are None values in the given params.", warn_only)


def unique_filename(data_type, data_source, extension='.pkl', folder=None):
    """<type>_<source>_<uuid4><extension> with blanks as dashes.  ref :94-100"""
    folder = example_storage if folder is None else Path(folder)
    return folder / "{}_{}_{}{}".format(data_type.replace(' ', '-'), data_source.replace(' ', '-'), uuid.uuid4(), extension)


def dict_to_yaml_key(d, pair_separator=' : ', item_separator=' | '):
    """"k1 : v1 | k2 : v2" over the sorted items.  ref :103-112"""
    return item_separator.join(pair_separator.join((str(k), str(v))) for k, v in sorted(d.items()))


def _read_catalog(retries=12, wait=5.):
    """the catalog as a dict ({} for an empty file); a scanner error (another job rewriting the file) is retried"""
    for attempt in range(retries):
        try:
            with open(example_catalog, 'r') as file:
                return yaml.safe_load(file) or {}
        except yaml.scanner.ScannerError:
            if attempt == retries - 1:
                raise
            time.sleep(wait)
    return {}


def new_cataloged_filename(data_type, data_source, params, extension='.pkl', overwrite_mode='delete'):
    """Register ``params`` under (data_type, data_source) and return the fresh filename; the file of an existing
    entry with the same parameters is deleted ('delete') or moved to the trash bin ('trash').  ref :115-185"""
    check_data_type(data_type)
    check_data_source(data_type, data_source)
    check_params(params)
    filename = unique_filename(data_type, data_source, extension)
    catalog = _read_catalog()
    entries = catalog.setdefault(data_type, {}).setdefault(data_source, {})
    if entries is None:
        entries = catalog[data_type][data_source] = {}
    key = dict_to_yaml_key(params)
    old = entries.get(key)
    if old is not None and Path(old['filename']).exists():
        print("Existing file with the given parameters found.")
        if overwrite_mode == 'delete':
            print("Overwriting old file.")
            Path(old['filename']).unlink()
        elif overwrite_mode == 'trash':
            print("Moving old file to %s." % example_trashbin)
            Path(old['filename']).rename(example_trashbin / Path(old['filename']).name)
    entry = {str(k): str(v) for k, v in params.items()}
    entry['filename'] = str(filename)
    entries[key] = entry
    with open(example_catalog, 'w') as file:
        yaml.safe_dump(catalog, file, width=float("inf"))
    return filename


def filename_from_catalog(data_type, data_source, params):
    """The filename registered for ``params``; FileNotFoundError if there is none.  ref :188-214"""
    check_data_type(data_type)
    check_data_source(data_type, data_source)
    check_params(params)
    catalog = _read_catalog(retries=1)
    key = dict_to_yaml_key(params)
    try:
        info = catalog[data_type][data_source].get(key)
    except (KeyError, AttributeError, TypeError) as error:
        warnings.warn("Ran into a KeyError: %s" % error)
        info = None
    if info is not None:
        return info['filename']
    raise FileNotFoundError("\nCould not find data type %s and data source %s with params %s in the catalog."
                            "\nThe associated key in the catalog would be:\n%s" % (data_type, data_source, params, key))
