"""Local run catalogue: stands in for the reference's PostgreSQL tables
``outputfile`` / ``modelimages`` / ``uvvsmodels`` (reference ``data/schema.sql:152-201``;
out of scope per SURVEY section 8).  Outputs are registered per Input key and kept
in memory; when ``NEXOCLOM_B200_SAVEPATH`` is set they are also pickled there, each with a
small ``.json`` sidecar (input key, idnum, npackets, totalsource), and the directory IS the
index: a later process finds the runs of an Input by the key prefix in the file names, the
way the reference finds them with its SQL query on the ``outputfile`` table
(``Input.search``, reference Input.py:121-172).
"""
import hashlib
import json
import os
import pickle

_outputs = {}       # filename -> Output
_by_key = {}        # input key -> [filename, ...]
_counter = [0]


def input_key(inputs):
    blob = json.dumps({g: getattr(inputs, g).as_dict()
                       for g in ('geometry', 'surfaceinteraction', 'forces', 'spatialdist',
                                 'speeddist', 'angulardist', 'options')},
                      sort_keys=True, default=str)
    return hashlib.sha1(blob.encode()).hexdigest()


def _savepath():
    return os.environ.get('NEXOCLOM_B200_SAVEPATH')


def _disk_idnums():
    """idnums of every run file in the save path (file names end in .<idnum>.pkl)."""
    sp = _savepath()
    ids = []
    if sp and os.path.isdir(sp):
        for name in os.listdir(sp):
            parts = name.split('.')
            if len(parts) == 3 and parts[2] == 'pkl' and parts[1].isdigit():
                ids.append(int(parts[1]))
    return ids


def _disk_entries(key):
    """(idnum, filename, npackets, totalsource) of the runs of this Input that are in the
    save path, from their sidecars -- no pickle is opened."""
    sp = _savepath()
    out = []
    if sp and os.path.isdir(sp):
        for name in sorted(os.listdir(sp)):
            if name.startswith(key[:12] + '.') and name.endswith('.pkl'):
                filename = os.path.join(sp, name)
                try:
                    with open(filename[:-4] + '.json') as f:
                        meta = json.load(f)
                except (OSError, ValueError):
                    continue
                if meta.get('key') == key:
                    out.append((int(meta['idnum']), filename, meta['npackets'],
                                meta['totalsource']))
    return out


def register(inputs, output):
    key = input_key(inputs)
    _counter[0] = max([_counter[0]] + _disk_idnums()) + 1
    idnum = _counter[0]
    savepath = _savepath()
    if savepath:
        os.makedirs(savepath, exist_ok=True)
        filename = os.path.join(savepath, f'{key[:12]}.{idnum:010d}.pkl')
    else:
        filename = f'mem://{key[:12]}/{idnum:010d}'
    output.idnum = idnum
    output.filename = filename
    _outputs[filename] = output
    _by_key.setdefault(key, []).append(filename)
    if savepath:
        with open(filename, 'wb') as f:
            pickle.dump(output, f, protocol=pickle.HIGHEST_PROTOCOL)
        with open(filename[:-4] + '.json', 'w') as f:
            json.dump({'key': key, 'idnum': idnum, 'npackets': int(output.npackets),
                       'totalsource': float(output.totalsource)}, f)
    return idnum, filename


def search(inputs):
    """(idnums, filenames, npackets, totalsource) of the runs of this Input: the ones this
    process made plus the ones earlier processes left in the save path."""
    key = input_key(inputs)
    found = {f: (_outputs[f].idnum, f, _outputs[f].npackets, _outputs[f].totalsource)
             for f in _by_key.get(key, [])}
    for entry in _disk_entries(key):
        found.setdefault(entry[1], entry)
    rows = sorted(found.values())
    return ([r[0] for r in rows], [r[1] for r in rows], sum(r[2] for r in rows),
            sum(r[3] for r in rows))


def fetch(filename):
    if filename in _outputs:
        return _outputs[filename]
    with open(filename, 'rb') as f:
        return pickle.load(f)


def delete(inputs, filename=None):
    key = input_key(inputs)
    files = list(_by_key.get(key, [])) + [e[1] for e in _disk_entries(key)]
    for f in dict.fromkeys(files):
        if filename is None or f == filename:
            if f in _by_key.get(key, []):
                _by_key[key].remove(f)
            _outputs.pop(f, None)
            for path in (f, f[:-4] + '.json'):
                if os.path.exists(path):
                    os.remove(path)
