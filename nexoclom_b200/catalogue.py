"""Local run catalogue: stands in for the reference's PostgreSQL tables
``outputfile`` / ``modelimages`` / ``uvvsmodels`` (reference ``data/schema.sql:152-201``;
out of scope per SURVEY section 8).  Outputs are registered per Input key, kept
in memory, and -- when ``NEXOCLOM_B200_SAVEPATH`` is set -- also pickled there.
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


def register(inputs, output):
    key = input_key(inputs)
    _counter[0] += 1
    idnum = _counter[0]
    savepath = os.environ.get('NEXOCLOM_B200_SAVEPATH')
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
    return idnum, filename


def search(inputs):
    key = input_key(inputs)
    files = list(_by_key.get(key, []))
    outs = [_outputs[f] for f in files]
    return ([o.idnum for o in outs], files, sum(o.npackets for o in outs),
            sum(o.totalsource for o in outs))


def fetch(filename):
    if filename in _outputs:
        return _outputs[filename]
    with open(filename, 'rb') as f:
        return pickle.load(f)


def delete(inputs, filename=None):
    key = input_key(inputs)
    for f in list(_by_key.get(key, [])):
        if filename is None or f == filename:
            _by_key[key].remove(f)
            _outputs.pop(f, None)
            if os.path.exists(f):
                os.remove(f)
