"""The reference's own CubeList tests (lib/mpdaf/obj/tests/test_cubelist.py) on mpdaf_b200.CubeList.

Same fixtures (three 5x4x3 cubes arange(60)*{0,1,5}, float32 on disk, var = 1), same
assertions: exact data, expmap == 3, header propagation, scales/offsets.
"""
import os

import numpy as np
import pytest
from numpy.testing import assert_array_equal

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
        assert_array_equal(cube.data, self.scaledcube)

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
