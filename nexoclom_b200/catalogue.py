"""Local run catalogue: stands in for the reference's PostgreSQL tables
``outputfile`` / ``modelimages`` / ``uvvsmodels`` (reference ``data/schema.sql:152-201``;
out of scope per SURVEY section 8).  Outputs are registered per Input key; when
``NEXOCLOM_B200_SAVEPATH`` is set they are also pickled there, each with a small ``.json``
sidecar (input key, idnum, npackets, totalsource), and the directory IS the index: a later
process finds the runs of an Input by the key prefix in the file names, the way the
reference finds them with its SQL query on the ``outputfile`` table (``Input.search``,
reference Input.py:121-172).

One process per GPU: several ranks may share a save path.  File names are claimed with
``O_CREAT | O_EXCL`` (no two processes can get the same idnum) and carry the rank, and a
rank's ``search`` only returns the files it wrote itself (the other ranks' packets reach the
products through the all-reduce, ``sharding.py``).

An Output registered here keeps its packets ON THE GPU (``Output._table``, a compacted
``PacketTable``) so that ModelImage / LOSResult in the same process never round-trip through
the host; the tables are kept under a byte budget (``NEXOCLOM_B200_RESIDENT_GB``, default 64)
with least-recently-used eviction to float32 host arrays, and an Output that was written to
disk drops its host-side DataFrames again (they are rebuilt on access).
"""
import hashlib
import json
import os
import pickle
from collections import OrderedDict

_outputs = {}       # filename -> Output
_by_key = {}        # input key -> [filename, ...]
_counter = [0]
_resident = OrderedDict()   # filename -> bytes on the GPU, least recently used first
_adopted = set()            # files written elsewhere (the reference's archive): never removed


def input_key(inputs):
    blob = json.dumps({g: getattr(inputs, g).as_dict()
                       for g in ('geometry', 'surfaceinteraction', 'forces', 'spatialdist',
                                 'speeddist', 'angulardist', 'options')},
                      sort_keys=True, default=str)
    return hashlib.sha1(blob.encode()).hexdigest()


def _savepath():
    return os.environ.get('NEXOCLOM_B200_SAVEPATH')


def _rank():
    from .sharding import rank_world
    return rank_world()[0]


def _parse(name):
    """(key12, idnum, rank) of a run file name  <key12>.<idnum>[.r<rank>].pkl, else None."""
    parts = name.split('.')
    if len(parts) in (3, 4) and parts[-1] == 'pkl' and parts[1].isdigit():
        rank = 0
        if len(parts) == 4:
            if not (parts[2].startswith('r') and parts[2][1:].isdigit()):
                return None
            rank = int(parts[2][1:])
        return parts[0], int(parts[1]), rank
    return None


def _disk_idnums():
    """idnums of every run file in the save path."""
    sp = _savepath()
    ids = []
    if sp and os.path.isdir(sp):
        for name in os.listdir(sp):
            p = _parse(name)
            if p:
                ids.append(p[1])
    return ids


def _disk_entries(key):
    """(idnum, filename, npackets, totalsource) of the runs of this Input that THIS rank left
    in the save path, from their sidecars -- no pickle is opened."""
    sp = _savepath()
    out = []
    rank = _rank()
    if sp and os.path.isdir(sp):
        for name in sorted(os.listdir(sp)):
            p = _parse(name)
            if p and p[0] == key[:12] and p[2] == rank:
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


def _claim(savepath, key):
    """Reserve the next free idnum by creating its file exclusively."""
    os.makedirs(savepath, exist_ok=True)
    rank = _rank()
    idnum = max([_counter[0]] + _disk_idnums()) + 1
    while True:
        suffix = f'.r{rank}' if rank else ''
        filename = os.path.join(savepath, f'{key[:12]}.{idnum:010d}{suffix}.pkl')
        taken = any(_parse(n) and _parse(n)[1] == idnum for n in os.listdir(savepath))
        if not taken:
            try:
                fd = os.open(filename, os.O_CREAT | os.O_EXCL | os.O_WRONLY)
                os.close(fd)
                return idnum, filename
            except FileExistsError:
                pass
        idnum += 1


def register(inputs, output):
    key = input_key(inputs)
    savepath = _savepath()
    if savepath:
        idnum, filename = _claim(savepath, key)
    else:
        idnum = _counter[0] + 1
        filename = f'mem://{key[:12]}/{idnum:010d}'
    _counter[0] = max(_counter[0], idnum)
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
        drop = getattr(output, '_drop_host_frames', None)
        if drop is not None:
            drop()                   # the file has them; they are rebuilt on access
    touch(output)
    return idnum, filename


def search(inputs):
    """(idnums, filenames, npackets, totalsource) of the runs of this Input made by this
    rank: the ones this process made plus the ones earlier processes left in the save path."""
    key = input_key(inputs)
    found = {f: (_outputs[f].idnum, f, _outputs[f].npackets, _outputs[f].totalsource)
             for f in _by_key.get(key, [])}
    for entry in _disk_entries(key):
        found.setdefault(entry[1], entry)
    rows = sorted(found.values())
    return ([r[0] for r in rows], [r[1] for r in rows], sum(r[2] for r in rows),
            sum(r[3] for r in rows))


def fetch(filename):
    """The Output registered under `filename` (the live object of this process, packets still
    on the GPU), or the one unpickled from disk (which is then kept and made resident on
    first use)."""
    if filename in _outputs:
        return _outputs[filename]
    from . import refpickle
    output = refpickle.load(filename)      # this package's files and the reference's alike
    _outputs[filename] = output
    return output


def adopt(filename):
    """Make a run file that was written elsewhere -- by the REFERENCE (its PostgreSQL
    `outputfile` table is not available here) or by another save path -- known to this
    process: it is read (``refpickle``), and registered under the key of its own inputs so
    that ``inputs.search()``, ``ModelImage`` and ``LOSResult`` find it.  Returns the Output."""
    output = fetch(filename)
    key = input_key(output.inputs)
    output.filename = filename
    _adopted.add(filename)
    if getattr(output, 'idnum', None) is None:
        _counter[0] += 1
        output.idnum = _counter[0]
    if filename not in _by_key.setdefault(key, []):
        _by_key[key].append(filename)
    return output


def delete(inputs, filename=None):
    key = input_key(inputs)
    files = list(_by_key.get(key, [])) + [e[1] for e in _disk_entries(key)]
    for f in dict.fromkeys(files):
        if filename is None or f == filename:
            if f in _by_key.get(key, []):
                _by_key[key].remove(f)
            out = _outputs.pop(f, None)
            if out is not None and getattr(out, '_release_device', None) is not None:
                out._release_device()
            _resident.pop(f, None)
            if f in _adopted:                  # someone else's archive: only forgotten
                _adopted.discard(f)
                continue
            for path in (f, f[:-4] + '.json'):
                if os.path.exists(path):
                    os.remove(path)


# ---- GPU residency budget ---------------------------------------------------------------
def _budget_bytes():
    return int(float(os.environ.get('NEXOCLOM_B200_RESIDENT_GB', '64')) * 2**30)


def touch(output):
    """Mark `output` as most recently used and evict other Outputs' device tables (to
    float32 host arrays) until the resident total is under the budget."""
    name = getattr(output, 'filename', None)
    nbytes = output._device_bytes() if getattr(output, '_device_bytes', None) else 0
    if name is None:
        return
    _resident.pop(name, None)
    if nbytes:
        _resident[name] = nbytes
    total = sum(_resident.values())
    budget = _budget_bytes()
    for victim in list(_resident):
        if total <= budget:
            break
        if victim == name:
            continue
        out = _outputs.get(victim)
        if out is not None:
            out._evict_to_host()
        total -= _resident.pop(victim)
