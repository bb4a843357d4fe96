"""The reference's own CubeList tests (lib/mpdaf/obj/tests/test_cubelist.py) on mpdaf_b200.CubeList.

Same fixtures (three 5x4x3 cubes arange(60)*{0,1,5}, float32 on disk, var = 1), same
assertions: exact data, expmap == 3, header propagation, scales/offsets.
"""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from mpdaf_b200 import CubeList
from mpdaf_b200 import fits as mfits


class TestCubeList:
    shape = (5, 4, 3)
    ncubes = 3
    cubevals = np.array([0, 1, 5])
    scalelist = [0.5, 1, 1.5]
    offsetlist = [1.5, 2, 2.5]

    @pytest.fixture(autouse=True)
    def _cubes(self, tmp_path):
        self.arr = np.arange(np.prod(self.shape)).reshape(self.shape)
        self.scaledcube = np.mean([(self.arr * val + self.offsetlist[k]) * self.scalelist[k]
                                   for k, val in enumerate(self.cubevals)], axis=0)
        self.combined_cube = np.mean([(self.arr * i) for i in self.cubevals], axis=0)
        self.cubenames = []
        for i in self.cubevals:
            filename = os.path.join(str(tmp_path), 'cube-%d.fits' % i)
            mfits.write_cube(filename, (self.arr * i).astype(np.float32), np.ones(self.shape, dtype=np.float32),
                             primary_header={'CUBEIDX': int(i), 'OBJECT': 'OBJECT %d' % i, 'EXPTIME': 100})
            self.cubenames.append(filename)
        self.expmap = np.full(self.shape, self.ncubes, dtype=int)

    def assert_header(self, cube):
        assert cube.primary_header['FOO'] == 'BAR'
        assert 'CUBEIDX' not in cube.primary_header
        assert cube.primary_header['OBJECT'] == 'OBJECT 0'
        assert cube.data_header['OBJECT'] == 'OBJECT 0'
        assert cube.primary_header['EXPTIME'] == 100 * self.ncubes

    def test_checks(self, tmp_path):
        other = os.path.join(str(tmp_path), 'cube-tmp.fits')
        mfits.write_cube(other, np.zeros((3, 2, 1), dtype=np.float32))
        clist = CubeList(self.cubenames[:1] + [other])
        assert clist.check_dim() is False
        assert CubeList(self.cubenames).check_dim() is True

    @pytest.mark.gpu
    def test_median(self):
        clist = CubeList(self.cubenames)
        cube, expmap, stat_pix = clist.median(header={'FOO': 'BAR'})
        self.assert_header(cube)
        assert_array_equal(cube.data, self.arr)
        assert_array_equal(expmap.data, self.expmap)
        assert stat_pix.colnames == ['FILENAME', 'NPIX_NAN']
        assert_array_equal(stat_pix['NPIX_NAN'], 0)

    @pytest.mark.gpu
    def test_combine(self):
        clist = CubeList(self.cubenames)
        cube, expmap, stat_pix = clist.combine(header={'FOO': 'BAR'})
        cube2, expmap2, stat_pix2 = clist.combine(header={'FOO': 'BAR'}, mad=True)
        assert_array_equal(cube.data, cube2.data)
        assert_array_equal(expmap.data, expmap2.data)

        self.assert_header(cube)
        assert_array_equal(cube.data, self.combined_cube)
        assert_array_equal(expmap.data, self.expmap)
        assert stat_pix.colnames == ['FILENAME', 'NPIX_NAN', 'NPIX_REJECTED']
        assert_array_equal(stat_pix['NPIX_REJECTED'], 0)
        assert cube.primary_header['NFILES'][0] == 3
        assert cube.primary_header['HIERARCH MPDAF METH1 ID'][0] == 'obj.cubelist.merging'

        cube = clist.combine(nclip=(5., 5.), var='stat_mean')[0]
        assert_array_equal(cube.data, self.combined_cube)

    @pytest.mark.gpu
    def test_combine_scale(self):
        clist = CubeList(self.cubenames, scalelist=self.scalelist, offsetlist=self.offsetlist)
        cube, expmap, stat_pix = clist.combine(header={'FOO': 'BAR'})
        # the reference's test asserts equality; the pair kernel's mean is pivot + sum(w - pivot) / n, a few ulp from
        # the exposure-order sum of w (contract: 1e-6 relative)
        assert_allclose(cube.data, self.scaledcube, rtol=1e-14, atol=0)

    @pytest.mark.gpu
    def test_pymedian(self):   # test_cubelist.py:132-138
        clist = CubeList(self.cubenames)
        cube, expmap, stat_pix = clist.pymedian(header={'FOO': 'BAR'})
        self.assert_header(cube)
        assert_array_equal(cube.data, self.arr)
        assert_array_equal(expmap.data, self.expmap)

    @pytest.mark.gpu
    def test_pycombine(self):   # test_cubelist.py:156-170
        clist = CubeList(self.cubenames)
        cube, expmap, stat_pix, rejmap = clist.pycombine(header={'FOO': 'BAR'})
        cube2, expmap2, stat_pix2, rejmap2 = clist.pycombine(header={'FOO': 'BAR'}, mad=True)
        assert_array_equal(cube.data, cube2.data)
        assert_array_equal(expmap.data, expmap2.data)
        self.assert_header(cube)
        assert_array_equal(cube.data, self.combined_cube)
        assert_array_equal(expmap.data, self.expmap)
        assert_array_equal(rejmap.data, 0)
        cube = clist.pycombine(nclip=(5., 5.), var='stat_mean')[0]
        assert_array_equal(cube.data, self.combined_cube)

    @pytest.mark.gpu
    def test_pycombine_scale(self):   # test_cubelist.py:179-184
        clist = CubeList(self.cubenames, scalelist=self.scalelist, offsetlist=self.offsetlist)
        cube2, expmap2, _, _ = clist.pycombine(header={'FOO': 'BAR'})
        assert_allclose(cube2.data, self.scaledcube, rtol=1e-14, atol=0)

    @pytest.mark.gpu
    def test_mosaic_combine(self):   # test_cubelist.py:186-197
        from mpdaf_b200 import CubeMosaic
        clist = CubeMosaic(self.cubenames, self.cubenames[0])
        cube, expmap, stat_pix, rejmap = clist.pycombine(header={'FOO': 'BAR'})
        self.assert_header(cube)
        assert_array_equal(cube.data, self.combined_cube)
        assert_array_equal(expmap.data, self.expmap)
        cube2, expmap2, _, _ = clist.pycombine(header={'FOO': 'BAR'}, mad=True)
        assert_array_equal(cube.data, cube2.data)
        assert_array_equal(expmap.data, expmap2.data)
        with pytest.raises(NotImplementedError):
            clist.combine()

    @pytest.mark.gpu
    def test_mosaic_offsets_and_rejmap(self, tmp_path):
        """Cubes placed at different CRPIX offsets inside a larger output grid; rejmap counts."""
        from mpdaf_b200 import CubeMosaic
        import oracle
        rng = np.random.default_rng(5)
        out_shape, sub = (6, 9, 11), (6, 5, 6)
        outfile = os.path.join(str(tmp_path), 'out.fits')
        mfits.write_cube(outfile, np.zeros(out_shape, dtype=np.float32), data_header={'CRPIX1': 6.0, 'CRPIX2': 5.0})
        offsets = [(0, 0), (4, 5), (2, 3), (1, 4)]
        names, padded, pvar = [], [], []
        for k, (oy, ox) in enumerate(offsets):
            d = (10 + rng.standard_normal(sub)).astype(np.float32)
            d[rng.random(sub) < 0.1] += 30
            v = rng.uniform(0.5, 2, sub).astype(np.float32)
            name = os.path.join(str(tmp_path), 'part-%d.fits' % k)
            mfits.write_cube(name, d, v, data_header={'CRPIX1': 6.0 - ox, 'CRPIX2': 5.0 - oy})
            names.append(name)
            full = np.full(out_shape, np.nan, dtype=np.float32)
            fv = np.full(out_shape, np.nan, dtype=np.float32)
            full[:, oy:oy + sub[1], ox:ox + sub[2]] = d
            fv[:, oy:oy + sub[1], ox:ox + sub[2]] = v
            padded.append(full)
            pvar.append(fv)
        cube, expmap, statpix, rejmap = CubeMosaic(names, outfile).pycombine(nclip=2.0)
        want = oracle.combine_port(padded, pvar, nclip=2.0)
        assert cube.data.shape == out_shape
        assert_array_equal(expmap.data.ravel(), want["expmap"])
        np.testing.assert_allclose(cube.data.ravel(), want["data"], rtol=1e-6, equal_nan=True)
        nvalid = sum((~np.isnan(p)).astype(int) for p in padded)
        assert_array_equal(rejmap.data, nvalid - expmap.data)
        assert (rejmap.data > 0).any() and (expmap.data == 0).any()

    @pytest.mark.gpu
    def test_from_arrays_and_write(self, tmp_path):
        arrays = [np.asarray(mfits.read_image(f, 'DATA'), dtype=np.float32) for f in self.cubenames]
        var = [np.ones(self.shape, dtype=np.float32)] * 3
        clist = CubeList.from_arrays(arrays, var)
        cube, expmap, stat_pix = clist.combine()
        assert_array_equal(cube.data, self.combined_cube)
        assert_array_equal(cube.var, np.full(self.shape, 3.0 / 9.0))
        out = os.path.join(str(tmp_path), 'combined.fits')
        cube.write(out)
        assert_array_equal(np.asarray(mfits.read_image(out, 'DATA')), self.combined_cube.astype(np.float32))
        assert mfits.read_hdus(out)[0][0]['NFILES'] == 3

    @pytest.mark.gpu
    def test_device_stack(self):
        from mpdaf_b200 import synth
        shape = (7, 6, 9)
        stack, _, _ = synth.device_stack(5, shape, seed=3)
        data, var = synth.host_stack(5, shape, seed=3)
        import oracle
        want = oracle.combine_port(data, var, nclip=3.0)
        cube, expmap, statpix = CubeList.from_arrays(stack).combine(nclip=3.0)
        assert_array_equal(expmap.data.ravel(), want["expmap"])
        np.testing.assert_allclose(cube.data.ravel(), want["data"], rtol=1e-6, equal_nan=True)
        assert_array_equal(statpix['NPIX_NAN'], int(np.prod(shape)) - want["valid_pix"])


@pytest.mark.gpu
def test_expnb_histogram_on_device():
    """_compute_expnb (cubelist.py:69-74) from the device-side exposure-map histogram."""
    from mpdaf_b200 import synth
    from mpdaf_b200.cubelist import _compute_expnb, _expnb_from_device
    data, var = synth.host_stack(7, (5, 12, 14), seed=9)
    clist = CubeList.from_arrays(data, var)
    cube, expmap, _ = clist.combine(nclip=2.0)
    assert _expnb_from_device(7) == _compute_expnb(expmap.data)
    from mpdaf_b200.ctools import ctools
    hist = np.zeros(8, dtype=np.int64)
    assert ctools.mpdaf_b200_last_expmap_histogram(hist, 8) == 0
    assert_array_equal(hist, np.bincount(expmap.data.ravel(), minlength=8))
    cube, expmap, _ = clist.median()
    assert _expnb_from_device(7) == _compute_expnb(expmap.data)


def _fits_standard_check(path):
    """Independent check of a written file against the FITS 4.0 standard -- written from the standard, not from
    mpdaf_b200/fits.py: 2880-byte blocks, 80-character ASCII cards, mandatory keywords in order, END card, data
    size |BITPIX| / 8 * prod(NAXISn) padded to a block, big-endian IEEE floats.  Returns {extname: array}."""
    raw = open(path, 'rb').read()
    assert len(raw) % 2880 == 0, 'file is not a whole number of 2880-byte blocks'
    pos, hdus, first = 0, {}, True
    while pos < len(raw):
        cards, done = [], False
        while not done:
            block = raw[pos:pos + 2880]
            assert len(block) == 2880
            pos += 2880
            for c in range(36):
                card = block[80 * c:80 * c + 80]
                assert all(32 <= b < 127 for b in card), 'header holds a non-printable byte'
                text = card.decode('ascii')
                if text.startswith('END') and text[3:].strip() == '':
                    assert all(x == 32 for x in block[80 * (c + 1):]), 'header block is not blank-padded after END'
                    done = True
                    break
                cards.append(text)
        def value(key):
            for text in cards:
                if text[:8].rstrip() == key and text[8:10] == '= ':
                    return text[10:].split('/')[0].strip()
            return None
        mandatory = [c[:8].rstrip() for c in cards[:3]]
        assert mandatory[0] == ('SIMPLE' if first else 'XTENSION') and mandatory[1:] == ['BITPIX', 'NAXIS'], mandatory
        if first:
            assert value('SIMPLE') == 'T'
        else:
            assert value('XTENSION').strip("'").strip() == 'IMAGE' and value('PCOUNT') == '0' and value('GCOUNT') == '1'
        bitpix, naxis = int(value('BITPIX')), int(value('NAXIS'))
        assert bitpix in (8, 16, 32, 64, -32, -64)
        axes = []
        for k in range(naxis):
            assert cards[3 + k][:8].rstrip() == 'NAXIS%d' % (k + 1), 'NAXISn cards must follow NAXIS in order'
            axes.append(int(value('NAXIS%d' % (k + 1))))
        nbytes = abs(bitpix) // 8 * int(np.prod(axes)) if naxis else 0
        data = raw[pos:pos + nbytes]
        assert len(data) == nbytes, 'data unit is truncated'
        padded = (nbytes + 2879) // 2880 * 2880
        assert all(b == 0 for b in raw[pos + nbytes:pos + padded]), 'data unit is not zero-padded to a block'
        pos += padded
        if naxis:
            dtype = {-32: '>f4', -64: '>f8', 32: '>i4', 16: '>i2', 8: 'u1', 64: '>i8'}[bitpix]
            name = (value('EXTNAME') or "'PRIMARY'").strip("'").strip()
            hdus[name] = np.frombuffer(data, dtype=dtype).reshape(axes[::-1])
        first = False
    return hdus


class TestWriterAndMosaicDetails:
    @pytest.mark.gpu
    def test_written_cube_is_standard_fits(self, tmp_path):
        """§8 f-4: the combined cube written by Cube.write is valid FITS for a reader that shares no code with the
        writer, and holds the float32 images of the result (NaN where nothing was combined)."""
        rng = np.random.default_rng(3)
        shape = (4, 7, 9)
        data = [(10 + rng.standard_normal(shape)).astype(np.float32) for _ in range(5)]
        var = [rng.uniform(0.5, 2, shape).astype(np.float32) for _ in range(5)]
        for d in data:
            d[0, 0, :] = np.nan
        from helpers import write_exposures
        files = write_exposures(tmp_path, data, var)
        cube, expmap, _ = CubeList(files).combine(nclip=3.0)
        out = str(tmp_path / 'combined.fits')
        cube.write(out)
        hdus = _fits_standard_check(out)
        assert set(hdus) >= {'DATA', 'STAT'}
        assert hdus['DATA'].shape == shape and hdus['DATA'].dtype == np.dtype('>f4')
        np.testing.assert_array_equal(hdus['DATA'].astype(np.float32), cube.data.astype(np.float32))
        np.testing.assert_array_equal(hdus['STAT'].astype(np.float32), cube.var.astype(np.float32))
        assert np.isnan(hdus['DATA'][0, 0]).all()
        for f in files:                                     # the inputs written by the same writer pass the same check
            _fits_standard_check(f)

    @pytest.mark.gpu
    def test_pycombine_streams_planes_and_writes_reference_keywords(self, tmp_path, monkeypatch):
        """§8 f-3: the py route places the exposures a few planes at a time (bounded host memory), counters and
        rejmap accumulate over the chunks, and the provenance keywords are the reference's (cubelist.py:199-205)."""
        import oracle
        rng = np.random.default_rng(4)
        shape = (9, 8, 10)
        n = 6
        data = [(10 + rng.standard_normal(shape)).astype(np.float32) for _ in range(n)]
        var = [rng.uniform(0.5, 2, shape).astype(np.float32) for _ in range(n)]
        for d in data:
            d[rng.random(shape) < 0.05] = np.nan
            d[rng.random(shape) < 0.03] += 30
        from helpers import write_exposures
        files = write_exposures(tmp_path, data, var)
        want = oracle.combine_port(data, var, nclip=3.0)
        whole = CubeList(files).pycombine(nclip=3.0)
        monkeypatch.setenv('MPDAF_B200_PYCOMBINE_BYTES', str(2 * n * 80 * 4 * 2))      # two planes per chunk
        parts = CubeList(files).pycombine(nclip=3.0)
        for cube, expmap, statpix, rejmap in (whole, parts):
            np.testing.assert_array_equal(expmap.data.ravel(), want['expmap'])
            np.testing.assert_allclose(cube.data.ravel(), want['data'], rtol=1e-6, equal_nan=True)
            np.testing.assert_allclose(cube.var.ravel(), want['var'], rtol=1e-6, equal_nan=True)
            np.testing.assert_array_equal(statpix['NPIX_REJECTED'], want['valid_pix'] - want['select_pix'])
            nvalid = sum((~np.isnan(d)).astype(np.int32) for d in data)
            np.testing.assert_array_equal(rejmap.data, nvalid - expmap.data)
            names = [cube.primary_header['HIERARCH MPDAF METH1 PARAM%d NAME' % k][0] for k in range(1, 6)]
            assert names == ['nmax', 'nclip_low', 'nclip_up', 'var', 'mad']
            assert cube.primary_header['HIERARCH MPDAF METH1 ID'][0] == 'obj.cubelist.pycombine'

    @pytest.mark.gpu
    def test_device_resident_cube_brings_its_finite_mask(self):
        """§8 f-1: for a device-resident stack the ~isfinite mask of the result is built on the device."""
        from mpdaf_b200 import synth
        shape, n = (5, 20, 24), 12
        stack, _, _ = synth.device_stack(n, shape, seed=3, device=0)
        cube, expmap, _ = CubeList.from_arrays(stack).combine(nclip=3.0)
        assert cube._mask is not None and cube._mask.dtype == bool and cube._mask.shape == cube.data.shape
        np.testing.assert_array_equal(cube.mask, ~np.isfinite(cube.data))
        assert cube.mask.any() and not cube.mask.all()
        med, _, _ = CubeList.from_arrays(stack).median()
        np.testing.assert_array_equal(med.mask, ~np.isfinite(med.data))
