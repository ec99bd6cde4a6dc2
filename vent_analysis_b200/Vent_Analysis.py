"""Drop-in for the reference's ``Vent_Analysis`` class (``/root/reference/Vent_Analysis.py``):
same constructor, attribute names, method names, metadata keys and output dtypes; the voxel
hot path -- ``calculate_SNR``, ``N4_bias_correction``, ``calculate_VDP``, ``calculateBorder``
and ``calculate_CI`` -- runs in the hand-written sm_100a kernels of ``libventb200`` through
its C ABI (``include/ventb200.h``).  There is no CPU fallback.

Additive to the reference: the ``vox=`` and ``device=`` constructor keywords (the reference's
array constructor crashes at Vent_Analysis.py:166 unless ``pickle_dict={'vox': ...}`` is used;
that route still works here).

Results are plain numpy arrays / scalars on ``self`` (nothing device-side is stored), so
``pickleMe`` / ``unPickleMe`` and the reference's exporters keep working.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

from . import CI, _lib


def _as_label_u8(A, what):
    A = np.asarray(A)
    if A.dtype == np.bool_ or A.dtype == np.uint8:
        return np.ascontiguousarray(A, dtype=np.uint8)
    a8 = A.astype(np.uint8)
    if not np.array_equal(a8, A):
        raise ValueError(f"{what}: expected an array of integer labels in [0, 255]")
    return np.ascontiguousarray(a8)


class Vent_Analysis:
    """Performs the complete VDP analysis: N4 bias correction, normalisation, defect
    calculation, VDP and cluster index (reference docstring: Vent_Analysis.py:27-57)."""

    def __init__(self, xenon_path=None, mask_path=None, proton_path=None, xenon_array=None, mask_array=None,
                 proton_array=None, pickle_dict=None, pickle_path=None, vox=None, device=0):
        self.version = '241007_vent'          # Vent_Analysis.py:67
        self.proton = ''
        self.N4HPvent = ''
        self.defectArray = ''
        self.CIarray = ''
        self.vox = ''
        self.ds = ''
        self.twix = ''
        self.raw_k = ''
        self.raw_HPvent = ''
        self.metadata = {k: '' for k in (
            'fileName', 'PatientName', 'PatientAge', 'PatientBirthDate', 'PatientSex', 'Disease', 'StudyDate',
            'SeriesTime', 'DE', 'SNR', 'VDP', 'VDP_lb', 'VDP_km', 'LungVolume', 'DefectVolume', 'CI', 'FEV1',
            'FVC', 'visit', 'IRB', 'treatment', 'analysisUser', 'notes')}      # :78-103
        self._device = int(device)
        if vox is not None:
            self.vox = [float(v) for v in vox]

        if xenon_array is not None:
            print(f'\033[34mXenon array provided: {xenon_array.shape}\033[37m')
            self.HPvent = xenon_array
        if xenon_path is not None:
            try:
                print('\033[34mXenon DICOM path provided. Opening DICOM...\033[37m')
                self.ds, self.HPvent = self.openSingleDICOM(xenon_path)
            except Exception:
                print('\033[31mOpening Xenon DICOM failed...\033[37m')
            try:
                print('\033[34mPulling Xenon DICOM Header\033[37m')
                self.pullDICOMHeader()
            except Exception:
                print('\033[31mPulling Xenon DICOM Header failed...\033[37m')

        if mask_array is not None:
            print(f'\033[34mMask array provided: {mask_array.shape}\033[37m')
            self.mask = mask_array
            self.mask_border = self.calculateBorder(self.mask)
        if mask_path is not None:
            try:
                print('\033[34mLoading Mask and calculating border\033[37m')
                _, self.mask = self.openDICOMfolder(mask_path)
                self.mask_border = self.calculateBorder(self.mask)
            except Exception:
                print('\033[31mLoading Mask and calculating border failed...\033[37m')

        if proton_array is not None:
            print(f'\033[34mProton array provided: {proton_array.shape}\033[37m')
            self.proton = proton_array
        if proton_path is not None:
            try:
                print('\033[34mProton DICOM Path provided. Opening...\033[37m')
                self.proton_ds, self.proton = self.openSingleDICOM(proton_path)
            except Exception:
                print('\033[31mOpening Proton DICOM failed...\033[37m')

        if pickle_path is not None:
            print(f'\033[34mPickle path provided: {pickle_path}. Loading...\033[37m')
            try:
                with open(pickle_path, 'rb') as file:
                    pickle_dict = pickle.load(file)
                print('\033[32mPickle file successfully loaded.\033[37m')
            except Exception:
                print('\033[31mOpening Pickle from path and building arrays failed...\033[37m')
        if pickle_dict is not None:
            self.unPickleMe(pickle_dict)

        # Vent_Analysis.py:166 (raises, like the reference, when vox is still unset)
        self.metadata['LungVolume'] = np.sum(self.mask == 1) * np.prod(np.divide(self.vox, 10)) / 1000

    # ------------------------------------------------------------------ file input (host side)
    def openSingleDICOM(self, dicom_path):
        """Vent_Analysis.py:169-181: one multi-frame DICOM -> [row, col, slice]."""
        import pydicom as dicom
        ds = dicom.dcmread(dicom_path, force=dicom_path is None)
        arr = np.transpose(ds.pixel_array, (1, 2, 0))
        print(f'\033[32mI opened a DICOM of shape {arr.shape}\033[37m')
        return ds, arr

    def openDICOMfolder(self, maskFolder):
        """Vent_Analysis.py:184-196: a folder of 2D DICOMs -> float64 [row, col, slice]."""
        import pydicom as dicom
        files = [f for f in sorted(os.listdir(maskFolder)) if f.endswith('.dcm')]
        ds = dicom.dcmread(os.path.join(maskFolder, files[0]))
        mask = np.zeros((ds.pixel_array.shape[0], ds.pixel_array.shape[1], len(files)))
        for k, f in enumerate(files):
            ds = dicom.dcmread(os.path.join(maskFolder, f))
            mask[:, :, k] = ds.pixel_array
        print(f'\033[32mI built a mask of shape {mask.shape}\033[37m')
        return ds, mask

    def pullDICOMHeader(self):
        """Vent_Analysis.py:198-223: patient fields + voxel size from the DICOM header."""
        for elem in ['PatientName', 'PatientAge', 'PatientBirthDate', 'PatientSize', 'PatientWeight', 'PatientSex',
                     'StudyDate', 'StudyTime', 'SeriesTime']:
            try:
                self.metadata[elem] = self.ds[elem].value
            except Exception:
                print(f'\033[31mNo {elem}\033[37m')
                self.metadata[elem] = ''
        for k in range(100):
            try:
                self.vox = self.ds[0x5200, 0x9230][k]['PixelMeasuresSequence'][0].PixelSpacing
                break
            except Exception:
                if k == 99:
                    print('Pixel Spacing not in correct place in DICOM header, please enter each dimension...')
                    self.vox = [float(input()), float(input())]
        try:
            self.vox = [float(self.vox[0]), float(self.vox[1]), float(self.ds.SpacingBetweenSlices)]
        except Exception:
            print('Slice spacing not in correct position in DICOM header. Please enter manually:')
            self.vox = [float(self.vox[0]), float(self.vox[1]), float(input())]
        self.metadata['LungVolume'] = np.sum(self.mask == 1) * np.prod(np.divide(self.vox, 10)) / 1000

    # --------------------------------------------------------------------------- GPU hot path
    def _ctx(self):
        return _lib.context(getattr(self, '_device', 0))

    def _mask_u8(self, mask=None):
        m = self.mask if mask is None else mask
        m8 = np.asarray(m).astype(np.float32).astype(np.uint8)     # Vent_Analysis.py:323, :327
        if not np.array_equal(m8 > 0, np.asarray(m) > 0) or m8.max(initial=0) > 1:
            raise ValueError("mask must be binary (0/1): the reference mixes mask==1, mask>0 and "
                             "sum(mask) (Vent_Analysis.py:166,:245,:251)")
        return np.ascontiguousarray(m8)

    def calculateBorder(self, A):
        """Vent_Analysis.py:225-231: per slice, np.gradient != 0 along rows or columns
        (float64 0/1).  Kernel: ``border_kernel``."""
        A = np.asarray(A)
        if A.ndim != 3:
            raise ValueError("calculateBorder expects a 3D array")
        if A.shape[0] < 2 or A.shape[1] < 2:
            raise ValueError("Shape of array too small to calculate a numerical gradient, "
                             "at least (edge_order + 1) elements are required.")
        a8 = _as_label_u8(A, "calculateBorder")
        out = self._ctx().pipeline_host([A.shape], _lib.DO_MASK_BORDER, mask=a8)
        return out["mask_border"].reshape(A.shape).astype(np.float64)

    def normalize(self, x):
        """Vent_Analysis.py:233-237."""
        if (np.max(x) - np.min(x)) == 0:
            return x
        return (x - np.min(x)) / (np.max(x) - np.min(x))

    def calculate_SNR(self, A, FOVbuffer=20, manualNoise=False):
        """Vent_Analysis.py:337-357.  ``FOVbuffer`` is overwritten with 20 by the reference
        (:343; ``calculate_VDP`` even passes the mask in that slot), so it is ignored here too.
        Kernel: ``snr_kernel`` (float64 accumulation); the result has the dtype numpy gives
        the reference (float32 for a float32 volume)."""
        A = np.asarray(A)
        out = self._ctx().pipeline_host([A.shape], _lib.DO_SNR, image=A.astype(np.float32), mask=self._mask_u8())
        st = out["stats"][0]
        if st["n_noise"] < 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
        snr = st["snr"]
        return np.float32(snr) if A.dtype == np.float32 else np.float64(snr)

    def N4_bias_correction(self, HPvent, mask):
        """Vent_Analysis.py:316-334 (SimpleITK N4BiasFieldCorrectionImageFilter, defaults) ->
        float32 array.  Kernel: ``n4_subject_kernel``."""
        HPvent = np.asarray(HPvent)
        out = self._ctx().pipeline_host([HPvent.shape], _lib.DO_N4, image=HPvent.astype(np.float32),
                                        mask=self._mask_u8(mask))
        self._n4_iters = [int(x) for x in out["stats"]["n4_iters"][0][:4]]
        return out["n4"].reshape(HPvent.shape)

    def calculate_VDP(self, thresh=0.6):
        """Vent_Analysis.py:239-263: SNR, N4, mean-anchored VDP (60 % threshold + per-slice 3x3
        median), defect border, linear-binning VDP -- one library call."""
        HP = np.asarray(self.HPvent)
        m8 = self._mask_u8()
        out = self._ctx().pipeline_host([HP.shape], _lib.DO_SNR | _lib.DO_N4 | _lib.DO_VDP,
                                        image=HP.astype(np.float32), mask=m8, thresh=thresh)
        st = out["stats"][0]
        if st["n_noise"] < 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
        if st["n_mask"] == 0:
            raise IndexError("list index out of range")           # signal_list[int(0*.99)] (:255)
        shape = HP.shape
        self.metadata['SNR'] = np.float32(st["snr"]) if HP.dtype == np.float32 else np.float64(st["snr"])
        self.N4HPvent = out["n4"].reshape(shape)
        self._n4_iters = [int(x) for x in st["n4_iters"][:4]]
        self.defectArray = out["defect"].reshape(shape).astype(np.float64)              # :247
        self.defectBorder = out["defect_border"].reshape(shape).astype(bool)            # :250
        nmask = np.sum(self.mask)
        self.metadata['VDP'] = 100 * np.float64(st["n_defect"]) / nmask                 # :251
        self.metadata['DefectVolume'] = np.int64(st["n_defect"]) * np.prod(np.divide(self.vox, 10)) / 1000  # :252
        self.defectArrayLB = out["defect_lb"].reshape(shape).astype(np.float64)         # :256
        self.metadata['VDP_lb'] = 100 * np.int64(st["n_lb12"]) / nmask                  # :257
        print('\033[32mcalculate_VDP ran successfully\033[37m')

    def calculate_CI(self):
        """Vent_Analysis.py:265-271: cluster-index array and its 95th-percentile value."""
        cv, r95 = CI._run(self.defectArray, self.vox, 50, device=getattr(self, '_device', 0))
        self.CIarray = cv * np.min(self.vox)                                            # CI.py:142
        self.metadata['CI'] = np.float64(r95 * np.min(self.vox))
        print(f"Calculated CI: {self.metadata['CI']}")

    # ----------------------------------------------------------------- consumers (host side)
    def build4DdataArray(self):
        """Vent_Analysis.py:292-313: Proton[0] HPvent[1] mask[2] N4HPvent[3] defectArray[4] CIarray[5]."""
        s = self.HPvent.shape
        dataArray = np.zeros((s[0], s[1], s[2], 6), dtype=np.float32)
        dataArray[:, :, :, 1] = self.HPvent
        dataArray[:, :, :, 2] = self.mask
        for ch, name in ((0, 'proton'), (3, 'N4HPvent'), (4, 'defectArray'), (5, 'CIarray')):
            try:
                dataArray[:, :, :, ch] = getattr(self, name)
            except Exception:
                print(f'\033[33m{name} does not exist (or has another shape) and was not added to 4D array\033[37m')
        return dataArray

    def cropToData(self, A, border=0, borderSlices=False):
        """Vent_Analysis.py:430-456: crop to the rows / cols / slices that hold signal."""
        A = np.asarray(A)
        rows = np.nonzero(A.sum(axis=(1, 2)) > 0)[0]
        cols = np.nonzero(A.sum(axis=(0, 2)) > 0)[0]
        slices = np.nonzero(A.sum(axis=(0, 1)) > 0)[0]
        sb = border if borderSlices else 0
        r0, r1 = max(rows[0] - border, 0), min(rows[-1] + border + 1, A.shape[0])
        c0, c1 = max(cols[0] - border, 0), min(cols[-1] + border + 1, A.shape[1])
        s0, s1 = max(slices[0] - sb, 0), min(slices[-1] + sb + 1, A.shape[2])
        return A[r0:r1, c0:c1, s0:s1], list(range(r0, r1)), list(range(c0, c1)), list(range(s0, s1))

    def pickleMe(self, pickle_path='C:/PIRL/data/VentPickle.pkl'):
        """Vent_Analysis.py:542-553."""
        pickle_dict = {}
        for attr in vars(self):
            try:
                pickle.dumps(getattr(self, attr))
                pickle_dict[attr] = getattr(self, attr)
            except (pickle.PicklingError, AttributeError, TypeError):
                print(f"\033[31mSkipping non-picklable attribute: {attr}\033[37m")
        with open(pickle_path, 'wb') as file:
            pickle.dump(pickle_dict, file)
        print(f'\033[32mPickled dictionary saved to {pickle_path}\033[37m')

    def unPickleMe(self, pickle_dict):
        """Vent_Analysis.py:556-559."""
        for attr, value in pickle_dict.items():
            setattr(self, attr, value)

    def _out_of_scope(self, name):
        raise NotImplementedError(
            f"{name} is a file-format exporter outside the N4 -> VDP -> CI hot path (SURVEY.md §2 rows 11-12); "
            "it consumes the arrays this class produces and is unchanged from the reference.")

    def exportNifti(self, filepath=None, fileName=None):
        self._out_of_scope("exportNifti")

    def exportDICOM(self, ds, save_dir='C:/PIRL/data/', optional_text='', forPACS=True):
        self._out_of_scope("exportDICOM")

    def screenShot(self, path='C:/PIRL/data/screenShotTest.png', normalize95=False):
        self._out_of_scope("screenShot")

    def process_RAW(self, filepath=None):
        self._out_of_scope("process_RAW")

    def dicom_to_json(self, ds, json_path='c:/pirl/data/DICOMjson.json', include_private=True):
        self._out_of_scope("dicom_to_json")

    def __repr__(self):
        """Vent_Analysis.py:562-577."""
        string = f'\033[35mVent_Analysis\033[37m class object version \033[94m{self.version}\033[37m\n'
        for attr, value in vars(self).items():
            if isinstance(value, str) and value == '':
                string += f'\033[31m {attr}: \033[37m\n'
            elif type(value) is np.ndarray:
                string += f'\033[32m {attr}: \033[36m{value.shape} \033[37m\n'
            elif type(value) is dict:
                for attr2, value2 in value.items():
                    if isinstance(value2, str) and value2 == '':
                        string += f'   \033[31m {attr2}: \033[37m\n'
                    else:
                        string += f'   \033[32m {attr2}: \033[36m{value2} \033[37m\n'
            else:
                string += f'\033[32m {attr}: \033[36m{type(value)} \033[37m\n'
        return string
