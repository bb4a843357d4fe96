"""Host side of the combine path: a `CubeList` with the reference's interface.

Mirrors lib/mpdaf/obj/cubelist.py of the reference for the two methods on the
hot path, `CubeList.median` (:432-468) and `CubeList.combine` (:470-536): same
arguments, same defaults (nmax=2, nclip=5.0, nstop=2, var='propagate',
mad=False), same output allocation and marshalling, the same
`(cube, expmap, statpix)` return, the same `statpix` arithmetic (:459-461,
:518-524), `_compute_expnb` (:69-74) and `save_combined_cube` header logic
(:392-430).  The work itself is done by the CUDA `_ctools` library through the
ctypes loader `mpdaf_b200.ctools` (the reference's `mpdaf.tools.ctools`).

astropy is not available in this image, so `Cube` here is a light container
(data / var ndarray, primary and data header dicts, `write`) rather than
`mpdaf.obj.Cube`, and `statpix` is a small column table; with astropy installed
`StatTable.to_astropy()` gives the reference's `astropy.table.Table`.

Besides FITS file names, a list can be built from in-memory exposure stacks
(`CubeList.from_arrays`), host NumPy arrays or device pointers, which is how the
FITS decoding stays outside the timed region in bench.py.
"""
import ctypes
import logging
import os
from ctypes import c_char_p

import numpy as np

from . import fits as _fits

__all__ = ('CubeList', 'CubeMosaic', 'Cube', 'StatTable', 'DeviceStack')

__version__ = '0.1'

# List of keywords copied to the combined cube (reference cubelist.py:52-66)
KEYWORDS_TO_COPY = (
    'ORIGIN', 'TELESCOP', 'INSTRUME', 'RA', 'DEC', 'EQUINOX', 'RADECSYS',
    'EXPTIME', 'MJD-OBS', 'DATE-OBS', 'PI-COI', 'OBSERVER', 'OBJECT',
    'HIERARCH ESO INS DROT POSANG',
    'HIERARCH ESO INS MODE',
    'HIERARCH ESO DET READ CURID',
    'HIERARCH ESO INS TEMP11 VAL',
    'HIERARCH ESO OBS ID',
    'HIERARCH ESO OBS NAME',
    'HIERARCH ESO OBS START',
    'HIERARCH ESO TEL AIRM END',
    'HIERARCH ESO TEL AIRM START',
    'HIERARCH ESO TEL AMBI FWHM END',
    'HIERARCH ESO TEL AMBI FWHM START'
)


def _compute_expnb(expmap):
    """Most frequent non-zero exposure count (reference cubelist.py:69-74).

    scipy.stats.mode returns the smallest of tied modes; the first maximum of a
    histogram does the same in one pass over the int32 map."""
    counts = np.bincount(np.asarray(expmap).ravel())
    if counts.size <= 1 or counts[1:].max() == 0:
        return 0
    return int(np.argmax(counts[1:])) + 1


def _expnb_from_device(nfiles):
    """expnb from the exposure-map histogram the library accumulated on the device during the
    merging call just made; same tie rule as `_compute_expnb`."""
    from .ctools import ctools
    hist = np.zeros(nfiles + 1, dtype=np.int64)
    if ctools.mpdaf_b200_last_expmap_histogram(hist, nfiles + 1) != 0 or hist[1:].max(initial=0) == 0:
        return None
    return int(np.argmax(hist[1:])) + 1


def _add_method_keywords(header, method, params, values, comments):
    """HIERARCH MPDAF METHn provenance block (reference tools/fits.py:54-89)."""
    i = 1
    while 'HIERARCH MPDAF METH%d ID' % i in header:
        i += 1
    header['HIERARCH MPDAF METH%d VERSION' % i] = (__version__, 'MPDAF version')
    header['HIERARCH MPDAF METH%d ID' % i] = (method, 'MPDAF method identifier')
    for p, (name, value, comment) in enumerate(zip(params, values, comments)):
        header['HIERARCH MPDAF METH%d PARAM%d NAME' % (i, p + 1)] = (name, comment)
        keyword = 'HIERARCH MPDAF METH%d PARAM%d VALUE' % (i, p + 1)
        if isinstance(value, str):
            value = value[0:80 - len(keyword) - 14]
        elif isinstance(value, (bool, np.bool_)):
            value = bool(value)
        header[keyword] = value


class Cube:
    """Combined cube: `data` (nz, ny, nx), optional `var`, two header dicts."""

    def __init__(self, data, var=None, primary_header=None, data_header=None, unit=None,
                 filename=None, comments=(), mask=None):
        self._mask = mask
        self.data = data
        self.var = var
        self.primary_header = dict(primary_header or {})
        self.data_header = dict(data_header or {})
        self.unit = unit
        self.filename = filename
        self.comments = list(comments)

    @property
    def shape(self):
        return self.data.shape

    @property
    def mask(self):
        """True where the voxel holds no finite value (reference Cube masks ~isfinite).  Device-resident
        combinations bring the mask with them (built on the device, `mpdaf_b200_finite_mask`)."""
        if self._mask is not None:
            return self._mask
        return ~np.isfinite(self.data)

    def write(self, filename, dtype=np.float32):
        """PRIMARY + DATA [+ STAT] float32 images, NaN for empty voxels (savemask='nan')."""
        _fits.write_cube(filename, self.data, self.var, self.primary_header, self.data_header,
                         dtype=dtype, comments=self.comments)
        self.filename = filename


class StatTable:
    """Column table standing in for astropy.table.Table (FILENAME, NPIX_NAN, ...)."""

    def __init__(self, columns, names):
        self.colnames = list(names)
        self._cols = {n: np.asarray(c) for n, c in zip(names, columns)}

    def __getitem__(self, name):
        return self._cols[name]

    def __len__(self):
        return len(next(iter(self._cols.values())))

    def to_astropy(self):
        from astropy.table import Table
        return Table([self._cols[n] for n in self.colnames], names=self.colnames)

    def __repr__(self):
        rows = ["  ".join(self.colnames)]
        for r in range(len(self)):
            rows.append("  ".join(str(self._cols[n][r]) for n in self.colnames))
        return "\n".join(rows)


class DeviceStack:
    """Exposure stack resident in device memory: raw pointers + geometry.

    data_ptrs / var_ptrs: one device address per exposure (nvox float32 or float64
    samples each, x fastest).  `keepalive` holds whatever owns the memory."""

    def __init__(self, data_ptrs, var_ptrs, shape, elem_type=4, device=0, keepalive=None):
        self.data_ptrs = [int(p) for p in data_ptrs]
        self.var_ptrs = None if var_ptrs is None else [int(p) for p in var_ptrs]
        self.shape = tuple(int(s) for s in shape)
        self.elem_type = int(elem_type)
        self.device = int(device)
        self.keepalive = keepalive


def _ptr_table(addresses):
    arr = (ctypes.c_void_p * len(addresses))(*addresses)
    return arr, ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))


class CubeList:

    """Manages a list of cubes and handles the combination.

    To run the combination, all the cubes must have the same dimensions and be
    on the same WCS grid. A global flux offset and scale can be given for each
    cube: ``(data + offset) * scale``.

    Parameters
    ----------
    files : list of str
        List of cubes FITS filenames.
    scalelist: list of float, optional
        List of scales to be applied to each cube.
    offsetlist: list of float, optional
        List of offsets to be applied to each cube.

    Attributes
    ----------
    files, nfiles, flux_scales, flux_offsets, shape : as in the reference
    (lib/mpdaf/obj/cubelist.py:268-303); wcs / wave objects are out of scope, the
    WCS keywords of cube 0's DATA header are carried over verbatim instead.
    """

    checkers = ('check_dim',)

    def __init__(self, files, scalelist=None, offsetlist=None):
        self._logger = logging.getLogger(__name__)
        self.files = list(files)
        self.nfiles = len(self.files)
        self._stack = None
        self._host_arrays = None

        # headers only, like DataArray's lazy loading (data.py:57-104)
        self._headers = []
        for f in self.files:
            hdus = _fits.read_hdus(f)
            data_hdr = next((h for h, _, _ in hdus if str(h.get('EXTNAME', '')).upper() == 'DATA'), None)
            if data_hdr is None:
                raise OSError(f'{f}: no DATA extension')
            self._headers.append((hdus[0][0], data_hdr))
        self._set_defaults()
        self.check_compatibility()

        self.flux_scales = scalelist
        self.flux_offsets = offsetlist

    #: devices the host-resident and file routes spread their wavelength slabs over: None = every usable B200 of the box
    #: (the library's default, MPDAF_B200_DEVICES narrows it), an int = that device only, a sequence = exactly those
    devices = None

    def _device(self):
        """`device` argument of a HOST-location call (include/mpdaf_b200.h): ALL_DEVICES or one index."""
        from .ctools import ALL_DEVICES, check, ctools
        if self.devices is None:
            return ALL_DEVICES
        if isinstance(self.devices, (int, np.integer)):
            return int(self.devices)
        ids = [int(d) for d in self.devices]
        check(ctools.mpdaf_b200_set_devices(len(ids), (ctypes.c_int * len(ids))(*ids)), 'mpdaf_b200_set_devices')
        return ALL_DEVICES

    @classmethod
    def from_arrays(cls, data, var=None, scalelist=None, offsetlist=None, names=None,
                    primary_header=None, data_header=None):
        """Build a list from in-memory exposures instead of FITS files.

        data / var: sequences of equally shaped (nz, ny, nx) float32 or float64
        C-contiguous NumPy arrays, or a `DeviceStack`."""
        self = cls.__new__(cls)
        self._logger = logging.getLogger(__name__)
        if isinstance(data, DeviceStack):
            self._stack = data
            self._host_arrays = None
            self.nfiles = len(data.data_ptrs)
            self.shape = tuple(data.shape)
        else:
            arrays = [np.ascontiguousarray(a) for a in data]
            dt = arrays[0].dtype
            if dt not in (np.float32, np.float64) or any(a.dtype != dt or a.shape != arrays[0].shape for a in arrays):
                raise ValueError('exposures must share one shape and be float32 or float64')
            vars_ = None
            if var is not None:
                vars_ = [np.ascontiguousarray(a, dtype=dt) for a in var]
                if len(vars_) != len(arrays) or any(a.shape != arrays[0].shape for a in vars_):
                    raise ValueError('var must match data')
            self._stack = None
            self._host_arrays = (arrays, vars_)
            self.nfiles = len(arrays)
            self.shape = tuple(arrays[0].shape)
        self.files = list(names) if names is not None else ['exposure-%03d' % i for i in range(self.nfiles)]
        self._headers = [(dict(primary_header or {}), dict(data_header or {}))] * self.nfiles
        self.unit = None
        self.flux_scales = scalelist
        self.flux_offsets = offsetlist
        return self

    def _set_defaults(self):
        h = self._headers[0][1]
        self.shape = (int(h['NAXIS3']), int(h['NAXIS2']), int(h['NAXIS1']))
        self.unit = h.get('BUNIT')

    def check_dim(self):
        """Checks if all cubes have same dimensions."""
        shapes = np.array([(int(h['NAXIS3']), int(h['NAXIS2']), int(h['NAXIS1'])) for _, h in self._headers])
        if not np.all(shapes == self.shape):
            self._logger.warning('all cubes have not same dimensions')
            for i in range(self.nfiles):
                self._logger.warning('%i X %i X %i cube (%s)', shapes[i, 0], shapes[i, 1], shapes[i, 2],
                                     self.files[i])
            return False
        return True

    def check_compatibility(self):
        """Checks if all cubes are compatible."""
        for checker in self.checkers:
            getattr(self, checker)()

    # ------------------------------------------------------------------
    def save_combined_cube(self, data, var=None, method='', keywords=None, expnb=None, unit=None,
                           header=None):
        if data.ndim != 3:
            data = data.reshape(self.shape)
        if var is not None and var.ndim != 3:
            var = var.reshape(self.shape)

        src_primary, src_data = self._headers[0]
        hdr = {}
        for key in KEYWORDS_TO_COPY:
            if key in src_primary:
                hdr[key] = src_primary[key]
        if expnb is not None and 'EXPTIME' in hdr:
            hdr['EXPTIME'] = hdr['EXPTIME'] * expnb
        if header is not None:
            hdr.update(header)

        data_header = {k: v for k, v in src_data.items()
                       if k.startswith(('CRPIX', 'CRVAL', 'CDELT', 'CD1_', 'CD2_', 'CD3_', 'CTYPE', 'CUNIT'))}
        if unit is None and self.unit is not None:
            data_header['BUNIT'] = self.unit
        if 'OBJECT' in hdr:
            data_header['OBJECT'] = hdr['OBJECT']
        elif 'OBJECT' in src_data:
            data_header['OBJECT'] = src_data['OBJECT']

        if keywords is not None:
            params, values, comments = list(zip(*keywords))
        else:
            params, values, comments = [], [], []
        _add_method_keywords(hdr, method, params, values, comments)
        hdr['NFILES'] = (len(self.files), 'number of files merged in the cube')

        comments_out = ['List of cubes merged in this cube:']
        comments_out += ['- ' + os.path.basename(f) for f in self.files]
        return Cube(data, var=var, primary_header=hdr, data_header=data_header, unit=unit or self.unit,
                    comments=comments_out)

    # ------------------------------------------------------------------
    def _scales_offsets(self):
        if self.flux_scales is None:
            scale = np.ones(self.nfiles, dtype=float)
        else:
            scale = np.asarray(self.flux_scales, dtype=float)
            self._logger.info('Using scales')
        if self.flux_offsets is None:
            offset = np.zeros(self.nfiles, dtype=float)
        else:
            offset = np.asarray(self.flux_offsets, dtype=float)
            self._logger.info('Using offsets')
        return scale, offset

    def median(self, header=None):
        """Combines cubes in a single data cube using median.

        Returns
        -------
        out : Cube, Cube, StatTable
            cube, expmap, statpix (columns FILENAME and NPIX_NAN)
        """
        from .ctools import ctools, check, HOST

        npixels = int(np.prod(self.shape))
        data = np.empty(npixels, dtype=np.float64, order='C')
        expmap = np.empty(npixels, dtype=np.intc, order='C')
        valid_pix = np.zeros(self.nfiles, dtype=np.intc, order='C')

        if self._host_arrays is None and self._stack is None:
            files = '\n'.join(self.files).encode('utf8')
            rc = ctools.mpdaf_merging_median(c_char_p(files), data, expmap, valid_pix)
            check(rc, 'mpdaf_merging_median')
        elif self._host_arrays is not None:
            arrays, _ = self._host_arrays
            keep, dptr = _ptr_table([a.ctypes.data for a in arrays])
            rc = ctools.mpdaf_b200_median_arrays(
                self.nfiles, dptr, arrays[0].dtype.itemsize, npixels, HOST, self._device(),
                data.ctypes.data, expmap.ctypes.data, valid_pix, None)
            check(rc, 'mpdaf_b200_median_arrays')
        else:
            data, _, expmap = self._run_device_stack(True, valid_pix, None)

        no_valid_pix = npixels - valid_pix
        stat_pix = StatTable([self.files, no_valid_pix], names=['FILENAME', 'NPIX_NAN'])

        expnb = _expnb_from_device(self.nfiles)
        kwargs = dict(expnb=_compute_expnb(expmap) if expnb is None else expnb,
                      method='obj.cubelist.median', header=header)
        expmap = self.save_combined_cube(expmap, unit='dimensionless', **kwargs)
        cube = self.save_combined_cube(data, **kwargs)
        cube._mask = self.__dict__.pop('_device_mask', None)
        if cube._mask is not None:
            cube._mask = cube._mask.reshape(cube.data.shape)
        return cube, expmap, stat_pix

    def combine(self, nmax=2, nclip=5.0, nstop=2, var='propagate', mad=False, header=None):
        """Combines cubes in a single data cube using sigma clipped mean.

        Same parameters and return as the reference (cubelist.py:219-254, :470-536):
        nmax, nclip (scalar or (low, up)), nstop, var in {'propagate', 'stat_mean',
        'stat_one'}, mad.  Returns cube, expmap, statpix (FILENAME, NPIX_NAN,
        NPIX_REJECTED).
        """
        from .ctools import ctools, check, HOST

        if np.isscalar(nclip):
            nclip_low = nclip
            nclip_up = nclip
        else:
            nclip_low = nclip[0]
            nclip_up = nclip[1]

        npixels = self.shape[0] * self.shape[1] * self.shape[2]
        data = np.empty(npixels, dtype=np.float64, order='C')
        vardata = np.empty(npixels, dtype=np.float64, order='C')
        expmap = np.empty(npixels, dtype=np.intc, order='C')
        valid_pix = np.zeros(self.nfiles, dtype=np.intc, order='C')
        select_pix = np.zeros(self.nfiles, dtype=np.intc, order='C')

        if var == 'propagate':
            var_mean = 0
        elif var == 'stat_mean':
            var_mean = 1
        else:
            var_mean = 2

        scale, offset = self._scales_offsets()
        params = (nmax, np.float64(nclip_low), np.float64(nclip_up), nstop, np.int32(var_mean), np.int32(mad))

        if self._host_arrays is None and self._stack is None:
            files = '\n'.join(self.files).encode('utf8')
            rc = ctools.mpdaf_merging_sigma_clipping(
                c_char_p(files), data, vardata, expmap, scale, offset, select_pix, valid_pix, *params)
            check(rc, 'mpdaf_merging_sigma_clipping')
        elif self._host_arrays is not None:
            arrays, vars_ = self._host_arrays
            if var_mean == 0 and vars_ is None:
                raise ValueError("var='propagate' needs the exposure variances")
            keep_d, dptr = _ptr_table([a.ctypes.data for a in arrays])
            keep_v, vptr = (None, None) if vars_ is None else _ptr_table([a.ctypes.data for a in vars_])
            rc = ctools.mpdaf_b200_sigma_clipping_arrays(
                self.nfiles, dptr, vptr, arrays[0].dtype.itemsize, npixels, HOST, self._device(),
                data.ctypes.data, vardata.ctypes.data, expmap.ctypes.data,
                scale.ctypes.data, offset.ctypes.data, select_pix, valid_pix,
                int(nmax), float(nclip_low), float(nclip_up), int(nstop), int(var_mean), int(mad), None)
            check(rc, 'mpdaf_b200_sigma_clipping_arrays')
        else:
            data, vardata, expmap = self._run_device_stack(
                False, valid_pix, select_pix, scale, offset,
                (int(nmax), float(nclip_low), float(nclip_up), int(nstop), int(var_mean), int(mad)))

        # no valid pixels
        with np.errstate(divide='ignore', invalid='ignore'):
            rej = (valid_pix - select_pix) / valid_pix.astype(float) * 100.0
        rej = " ".join(f"{p:.2f}%" for p in rej)
        self._logger.info("%% of rejected pixels per files: %s", rej)
        no_valid_pix = npixels - valid_pix
        rejected_pix = valid_pix - select_pix
        statpix = StatTable([self.files, no_valid_pix, rejected_pix],
                            names=['FILENAME', 'NPIX_NAN', 'NPIX_REJECTED'])

        keywords = [('nmax', nmax, 'max number of clipping iterations'),
                    ('nclip_low', nclip_low, 'lower clipping parameter'),
                    ('nclip_up', nclip_up, 'upper clipping parameter'),
                    ('nstop', nstop, 'clipping minimum number'),
                    ('var', var, 'type of variance')]
        expnb = _expnb_from_device(self.nfiles)
        kwargs = dict(expnb=_compute_expnb(expmap) if expnb is None else expnb, keywords=keywords, header=header,
                      method='obj.cubelist.merging')
        expmap = self.save_combined_cube(expmap, unit='dimensionless', **kwargs)
        cube = self.save_combined_cube(data, var=vardata, **kwargs)
        cube._mask = self.__dict__.pop('_device_mask', None)
        if cube._mask is not None:
            cube._mask = cube._mask.reshape(cube.data.shape)
        return cube, expmap, statpix

    # ------------------------------------------------------------------
    # The reference's "py" route (cubelist.py:77-216, :538-576; merging.pyx:61-125): same statistics,
    # optional per-cube (y, x) placement inside a larger output grid, and a `rejmap` output.  Here it
    # is the same CUDA path fed with NaN-padded exposures; `rejmap` = valid samples - expmap.
    def _placed_planes(self, images, z0, z1, pos, shapes):
        """Planes [z0, z1) of every exposure on the output grid (NaN outside an exposure's footprint)."""
        nz, ny, nx = self.shape
        out = []
        for i, img in enumerate(images):
            if pos is None and img.shape[1:] == (ny, nx):
                out.append(np.ascontiguousarray(img[z0:z1], dtype=np.float32))
                continue
            arr = np.full((z1 - z0, ny, nx), np.nan, dtype=np.float32)
            y0, x0 = (0, 0) if pos is None else (int(pos[i][0]), int(pos[i][1]))
            sy, sx = img.shape[1:] if shapes is None else (int(shapes[i][0]), int(shapes[i][1]))
            ys, xs = max(y0, 0), max(x0, 0)
            ye, xe = min(y0 + sy, ny), min(x0 + sx, nx)
            if ye > ys and xe > xs:
                arr[:, ys:ye, xs:xe] = img[z0:z1, ys - y0:ye - y0, xs - x0:xe - x0]
            out.append(arr)
        return out

    def _pycombine(self, nmax=2, nclip=5.0, var='propagate', nstop=2, nl=None, header=None, mad=False,
                   pos=None, shapes=None, method='obj.cubelist.pycombine'):
        """The reference's "py" route (cubelist.py:77-216): like the reference, which walks the cubes plane by plane
        (cubelist.py:141-152), the exposures are placed on the output grid a few wavelength planes at a time -- host
        memory stays bounded (MPDAF_B200_PYCOMBINE_BYTES, default 1 GiB of placed samples) however large the mosaic
        -- and every chunk goes through the same CUDA path; the counters accumulate over the chunks."""
        import os
        from .ctools import ctools, check, HOST
        if var not in ('propagate', 'stat_mean', 'stat_one'):
            raise ValueError("var must be 'propagate', 'stat_mean' or 'stat_one'")
        var_mean = {'propagate': 0, 'stat_mean': 1, 'stat_one': 2}[var]
        nclip_low, nclip_up = (nclip, nclip) if np.isscalar(nclip) else nclip
        nz, ny, nx = self.shape
        nl = nz if nl is None else int(nl)
        n = self.nfiles
        plane = ny * nx
        images = [_fits.read_image(f, 'DATA') for f in self.files]
        stats = [_fits.read_image(f, 'STAT') for f in self.files] if var_mean == 0 else None
        budget = int(os.environ.get('MPDAF_B200_PYCOMBINE_BYTES', 1 << 30))
        step = max(1, min(nl, budget // max(1, n * plane * 4 * (2 if var_mean == 0 else 1))))
        data = np.empty(nl * plane, dtype=np.float64)
        vardata = np.empty(nl * plane, dtype=np.float64)
        expmap = np.empty(nl * plane, dtype=np.int32)
        rejmap_data = np.empty((nl, ny, nx), dtype=np.int32)
        valid_pix = np.zeros(n, dtype=np.int32)
        select_pix = np.zeros(n, dtype=np.int32)
        scale = np.ascontiguousarray(self.flux_scales if self.flux_scales is not None else np.ones(n), dtype=np.float64)
        offset = np.ascontiguousarray(self.flux_offsets if self.flux_offsets is not None else np.zeros(n), dtype=np.float64)
        hist = np.zeros(n + 1, dtype=np.int64)
        part = np.zeros(n + 1, dtype=np.int64)
        for z0 in range(0, nl, step):
            z1 = min(nl, z0 + step)
            d = self._placed_planes(images, z0, z1, pos, shapes)
            v = self._placed_planes(stats, z0, z1, pos, shapes) if stats is not None else None
            keep_d, dptr = _ptr_table([a.ctypes.data for a in d])
            keep_v, vptr = (None, None) if v is None else _ptr_table([a.ctypes.data for a in v])
            sl = slice(z0 * plane, z1 * plane)
            rc = ctools.mpdaf_b200_sigma_clipping_arrays(
                n, dptr, vptr, 4, (z1 - z0) * plane, HOST, self._device(), data[sl].ctypes.data, vardata[sl].ctypes.data,
                expmap[sl].ctypes.data, scale.ctypes.data, offset.ctypes.data, select_pix, valid_pix,
                int(nmax), float(nclip_low), float(nclip_up), int(nstop), int(var_mean), int(mad), None)
            check(rc, 'mpdaf_b200_sigma_clipping_arrays')
            if ctools.mpdaf_b200_last_expmap_histogram(part, n + 1) == 0:
                hist += part
            nvalid = np.zeros((z1 - z0, ny, nx), dtype=np.int32)
            for a in d:
                nvalid += ~np.isnan(a)
            rejmap_data[z0:z1] = nvalid - expmap[sl].reshape(z1 - z0, ny, nx)     # merging.pyx:110
        shape = (nl, ny, nx)
        data, vardata, expmap = data.reshape(shape), vardata.reshape(shape), expmap.reshape(shape)
        npixels = nl * plane
        statpix = StatTable([self.files, npixels - valid_pix, valid_pix - select_pix],
                            names=['FILENAME', 'NPIX_NAN', 'NPIX_REJECTED'])
        # provenance keywords of the reference's pycombine (cubelist.py:199-205)
        keywords = [('nmax', nmax, 'max number of clipping iterations'),
                    ('nclip_low', nclip_low, 'lower clipping parameter'),
                    ('nclip_up', nclip_up, 'upper clipping parameter'),
                    ('var', var, 'type of variance'),
                    ('mad', bool(mad), 'use of MAD')]
        expnb = int(np.argmax(hist[1:]) + 1) if hist[1:].any() else 0
        saved_shape = self.shape
        self.shape = shape
        try:
            kwargs = dict(expnb=expnb, keywords=keywords, header=header, method=method)
            expmap_cube = self.save_combined_cube(expmap, unit='dimensionless', **kwargs)
            cube = self.save_combined_cube(data, var=vardata, **kwargs)
        finally:
            self.shape = saved_shape
        rejmap = Cube(rejmap_data, primary_header=expmap_cube.primary_header, data_header=expmap_cube.data_header,
                      unit='dimensionless', comments=expmap_cube.comments)
        return cube, expmap_cube, statpix, rejmap

    def pycombine(self, nmax=2, nclip=5.0, var='propagate', nstop=2, nl=None, header=None, mad=False):
        """`CubeList.combine` with the reference's pycombine signature and 4-tuple return
        (cube, expmap, statpix, rejmap) (cubelist.py:572-576)."""
        return self._pycombine(nmax=nmax, nclip=nclip, var=var, nstop=nstop, nl=nl, header=header, mad=mad)

    def pymedian(self, header=None):
        """Same result as `median` (the reference's python-fitsio route, cubelist.py:538-570)."""
        return self.median(header=header)

    # ------------------------------------------------------------------
    def _run_device_stack(self, median, valid_pix, select_pix, scale=None, offset=None, params=None):
        """Device-resident stack: outputs are produced on the device and read back."""
        import torch
        from .ctools import ctools, check, DEVICE

        st = self._stack
        npixels = int(np.prod(st.shape))
        dev = torch.device('cuda', st.device)
        out = torch.empty(npixels, dtype=torch.float64, device=dev)
        expmap = torch.empty(npixels, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        keep_d, dptr = _ptr_table(st.data_ptrs)
        if median:
            rc = ctools.mpdaf_b200_median_arrays(
                self.nfiles, dptr, st.elem_type, npixels, DEVICE, st.device,
                out.data_ptr(), expmap.data_ptr(), valid_pix, stream)
            check(rc, 'mpdaf_b200_median_arrays')
            self._device_mask = self._finite_mask_on_device(out, npixels, st.device, stream)
            return out.cpu().numpy(), None, expmap.cpu().numpy()
        outvar = torch.empty(npixels, dtype=torch.float64, device=dev)
        keep_v, vptr = (None, None) if st.var_ptrs is None else _ptr_table(st.var_ptrs)
        rc = ctools.mpdaf_b200_sigma_clipping_arrays(
            self.nfiles, dptr, vptr, st.elem_type, npixels, DEVICE, st.device,
            out.data_ptr(), outvar.data_ptr(), expmap.data_ptr(),
            scale.ctypes.data, offset.ctypes.data, select_pix, valid_pix, *params, stream)
        check(rc, 'mpdaf_b200_sigma_clipping_arrays')
        self._device_mask = self._finite_mask_on_device(out, npixels, st.device, stream)
        return out.cpu().numpy(), outvar.cpu().numpy(), expmap.cpu().numpy()

    @staticmethod
    def _finite_mask_on_device(out, npixels, device, stream):
        """~isfinite of the combined cube, computed where the cube lies."""
        import torch
        from .ctools import ctools, check
        mask = torch.empty(max(npixels, 1), dtype=torch.uint8, device=out.device)
        nmasked = ctypes.c_longlong(0)
        rc = ctools.mpdaf_b200_finite_mask(out.data_ptr(), npixels, device, mask.data_ptr(), ctypes.byref(nmasked), stream)
        check(rc, 'mpdaf_b200_finite_mask')
        return mask[:npixels].cpu().numpy().astype(bool)


class CubeMosaic(CubeList):

    """Manages a list of cubes and handles the combination to make a mosaic.

    Same contract as the reference (cubelist.py:582-705): all cubes share one WCS grid, the
    ``CRPIX`` keywords give each cube's integer offset inside the output grid, whose shape comes
    from the ``output_wcs`` FITS cube.  Only `pycombine` is available.
    """

    def __init__(self, files, output_wcs, **kwargs):
        hdus = _fits.read_hdus(output_wcs)
        self._out_header = next(h for h, _, _ in hdus if str(h.get('EXTNAME', '')).upper() == 'DATA')
        super().__init__(files, **kwargs)

    def _set_defaults(self):
        h = self._out_header
        self.shape = (int(h['NAXIS3']), int(h['NAXIS2']), int(h['NAXIS1']))
        self.unit = h.get('BUNIT')

    def check_dim(self):
        """Checks if all cubes have the same spectral range."""
        nz = {int(h['NAXIS3']) for _, h in self._headers}
        assert len(nz) == 1, 'Cubes must have the same spectral range.'
        return True

    def __getitem__(self, item):
        raise ValueError('Operation forbidden')

    def combine(self):
        """This method is not implemented for CubeMosaic."""
        raise NotImplementedError

    def median(self):
        """This method is not implemented for CubeMosaic."""
        raise NotImplementedError

    def pymedian(self):
        """This method is not implemented for CubeMosaic."""
        raise NotImplementedError

    def pycombine(self, nmax=2, nclip=5.0, var='propagate', nstop=2, nl=None, header=None, mad=False):
        def crpix(h):   # (CRPIX2, CRPIX1): python (y, x) order, cubelist.py:697-699
            return np.array([float(h.get('CRPIX2', 1.0)), float(h.get('CRPIX1', 1.0))])
        out = crpix(self._out_header)
        pos = np.array([np.rint(out - crpix(h)) for _, h in self._headers], dtype=int)
        shapes = np.array([(int(h['NAXIS2']), int(h['NAXIS1'])) for _, h in self._headers])
        return self._pycombine(nmax=nmax, nclip=nclip, var=var, nstop=nstop, nl=nl, header=header, mad=mad,
                               pos=pos, shapes=shapes, method='obj.cubemosaic.pycombine')
