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
    # Out of the hot path: thin pydicom glue with the reference's signatures and
    # array conventions, so that the GUI's calls keep working.
    def openSingleDICOM(self, dicom_path):
        """One multi-frame DICOM -> (dataset, [row, col, slice] array)  (Vent_Analysis.py:169-181)."""
        import pydicom
        ds = pydicom.dcmread(dicom_path)
        return ds, np.moveaxis(ds.pixel_array, 0, -1)

    def openDICOMfolder(self, maskFolder):
        """A folder of single-slice DICOMs -> (last dataset, float64 [row, col, slice])  (:184-196)."""
        import pydicom
        names = sorted(f for f in os.listdir(maskFolder) if f.endswith('.dcm'))
        sets = [pydicom.dcmread(os.path.join(maskFolder, f)) for f in names]
        return sets[-1], np.stack([d.pixel_array for d in sets], axis=-1).astype(np.float64)

    def pullDICOMHeader(self):
        """Patient fields and the voxel size [mm] from ``self.ds``  (:198-223)."""
        for key in ('PatientName', 'PatientAge', 'PatientBirthDate', 'PatientSize', 'PatientWeight', 'PatientSex',
                    'StudyDate', 'StudyTime', 'SeriesTime'):
            self.metadata[key] = getattr(self.ds, key, '')
        inplane = None
        for frame in self.ds.get((0x5200, 0x9230), []):               # per-frame functional groups
            try:
                inplane = frame.PixelMeasuresSequence[0].PixelSpacing
                break
            except (AttributeError, IndexError):
                continue
        if inplane is None:
            raise ValueError("PixelSpacing not found in the DICOM header: pass vox=[dx, dy, dz] instead")
        self.vox = [float(inplane[0]), float(inplane[1]), float(self.ds.SpacingBetweenSlices)]
        self.metadata['LungVolume'] = np.sum(self.mask == 1) * np.prod(np.divide(self.vox, 10)) / 1000

    # --------------------------------------------------------------------------- GPU hot path
    def _ctx(self):
        return _lib.context(getattr(self, '_device', 0))

    def _mask_u8(self, mask=None, binary=False):
        """The mask as label bytes (Vent_Analysis.py:323, :327: float32 -> UInt8).  Like the reference,
        N4 works on label 1 and SNR on ``mask > 0`` whatever other labels are present; ``binary``:
        ``calculate_VDP`` needs a 0/1 mask (with other labels the reference's threshold image and
        linear-binning codes get multiplied by the label values, :249, :256 -- not reproduced)."""
        m = self.mask if mask is None else mask
        m8 = np.asarray(m).astype(np.float32).astype(np.uint8)
        if not np.array_equal(m8 > 0, np.asarray(m) > 0):
            raise ValueError("mask labels must be integers in [0, 255]")
        if binary and m8.max(initial=0) > 1:
            raise ValueError("calculate_VDP needs a binary (0/1) mask: with other labels the reference weights its "
                             "threshold image by the label (Vent_Analysis.py:249,:256)")
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
        m8 = self._mask_u8(binary=True)
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
        """Every picklable attribute as one dict in a pickle file  (Vent_Analysis.py:542-553)."""
        state = {}
        for name, value in vars(self).items():
            try:
                pickle.dumps(value)
            except Exception:
                print(f"\033[31mSkipping non-picklable attribute: {name}\033[37m")
                continue
            state[name] = value
        with open(pickle_path, 'wb') as f:
            pickle.dump(state, f)
        print(f'\033[32mPickled dictionary saved to {pickle_path}\033[37m')

    def unPickleMe(self, pickle_dict):
        """Restore the attributes saved by :meth:`pickleMe`  (:556-559)."""
        vars(self).update(pickle_dict)

    # ---------------------------------------------------- overlay compositions
    def _overlay_rgb(self):
        """uint8 [rows, cols, slices, 3]: defects in red on the normalised |N4HPvent|
        (Vent_Analysis.py:387-391).  Kernels: ``ovl_minmax_kernel`` + ``ovl_rgb_kernel``."""
        shape = self.N4HPvent.shape
        d8 = np.ascontiguousarray(np.asarray(self.defectArray) == 1, dtype=np.uint8)
        return self._ctx().overlay_rgb_host([shape], self.N4HPvent, d8).reshape(shape + (3,))

    def exportDICOM(self, ds, save_dir='C:/PIRL/data/', optional_text='', forPACS=True):
        """Vent_Analysis.py:380-428: the defect overlay as RGB DICOM(s).  The pixel arithmetic runs on
        the GPU; writing the files is pydicom's."""
        if self.metadata['VDP'] == '':
            print('\033[31mCant export dicoms until you run calculate_VDP()...\033[37m')
            return
        import pydicom
        rgb = self._overlay_rgb()
        rows, cols, slices = rgb.shape[:3]
        desc = f"{optional_text} - VDP: {np.round(self.metadata['VDP'], 1)}"
        ds.PhotometricInterpretation, ds.SamplesPerPixel = 'RGB', 3
        ds.BitsAllocated = ds.BitsStored = 8
        ds.SeriesDescription = desc
        if not forPACS:                                   # one multi-frame object, slices first
            ds.Rows, ds.Columns, ds.NumberOfFrames = rows, cols, slices
            ds.HighBit = 7
            ds.SOPInstanceUID = ds.SeriesInstanceUID = pydicom.uid.generate_uid()
            ds.PixelData = np.ascontiguousarray(np.moveaxis(rgb, 2, 0)).tobytes()
            path = os.path.join(save_dir, f"{self.metadata['PatientName']}_defectDICOM.dcm")
            ds.save_as(path)
            print(f'\033[32mdefect DICOM saved to {path}\033[37m')
            return
        self.ds.SeriesInstanceUID = pydicom.uid.generate_uid()
        folder = os.path.join(save_dir, 'defectDICOMS')
        os.makedirs(folder, exist_ok=True)
        for k in range(slices):                           # one file per slice
            ds.PixelData = np.ascontiguousarray(rgb[:, :, k, :]).tobytes()
            ds.Rows, ds.Columns, ds.NumberOfFrames = rows, cols, 1
            ds.InstanceNumber, ds.SliceLocation = k + 1, k
            ds.SOPInstanceUID = pydicom.uid.generate_uid()
            ds.save_as(os.path.join(folder, f"dicom_{k}.dcm"))

    def _parula(self, parula=None):
        if parula is not None:
            return np.asarray(parula, dtype=np.float64)
        for path in (r'C:\PIRL\data\parula.np.npy', 'parula.npy', os.path.join(os.path.dirname(__file__), 'parula.npy')):
            if os.path.exists(path):
                return np.load(path)
        raise FileNotFoundError("parula colour table (64 x 3) not found: pass parula= or put parula.npy in the working directory")

    def screenShotArray(self, parula=None):
        """The montage image of Vent_Analysis.py:466-494 before any text: uint8 [7 R, S C, 3] (bands:
        blank, blank, proton, HPvent, N4 + mask border, N4 + defects, N4 + CI in parula colours).
        Kernels: ``ovl_cropbox_kernel``, ``ovl_minmax_kernel``, ``ovl_montage_kernel``."""
        ctx, shape = self._ctx(), np.asarray(self.HPvent).shape
        m8 = self._mask_u8()
        kw = {}
        if not isinstance(self.CIarray, str):
            from .sphere_table import get_sphere_pix
            vox = [float(v) for v in self.vox]
            tbl = get_sphere_pix(vox, 50)
            h = ctx.ci_table(tbl, shape, key=(tuple(vox), 50, tbl.shape[0]))
            radii = ctx.ci_radii(h)
            shell = np.full(shape, _lib.CI_OFF, dtype=np.uint16)
            sel = np.asarray(self.CIarray) > 0
            shell[sel] = np.searchsorted(radii * np.min(self.vox), np.asarray(self.CIarray)[sel])
            kw = dict(ci_shell=shell, handles=[h], vox_min=[float(np.min(self.vox))])
        imgs, _ = ctx.montage_host([shape], self.HPvent, self.N4HPvent, m8, np.asarray(self.mask_border) != 0,
                                   np.asarray(self.defectArray) > 0, self._parula(parula), proton=self.proton, **kw)
        return imgs[0]

    def screenShot(self, path='C:/PIRL/data/screenShotTest.png', normalize95=False, parula=None):
        """Vent_Analysis.py:458-520: the montage with the study text, saved as PNG."""
        from PIL import Image, ImageDraw, ImageFont
        image = Image.fromarray(self.screenShotArray(parula))
        try:
            font = ImageFont.truetype('arial.ttf', size=35)
        except OSError:
            font = ImageFont.load_default()
        draw, md = ImageDraw.Draw(image), self.metadata
        h = image.size[1] / 7
        lines = [(0.00, 0.10, f"Patient: {md['PatientName']} ({md['PatientAge']}/{md['PatientSex']})"),
                 (0.00, 0.40, f"Disease: {md['Disease']}"), (0.00, 0.70, f"StudyDate: {md['StudyDate']}"),
                 (0.00, 1.00, f"Visit#: {md['visit']}"), (0.00, 1.30, f"Treatment: {md['treatment']}"),
                 (0.25, 0.10, f"Lung Volume: {np.round(md['LungVolume'] * 1000)} mL"),
                 (0.25, 0.40, f"Defect Volume: {np.round(md['DefectVolume'] * 1000)} mL" if md['DefectVolume'] != '' else ''),
                 (0.50, 0.10, f"DE: {md['DE']} mL"), (0.50, 0.40, f"FEV1: {md['FEV1']} %"),
                 (0.50, 0.70, f"VDP: {np.round(md['VDP'], 1)} %" if md['VDP'] != '' else ''),
                 (0.50, 1.00, f"CI: {np.round(md['CI'])} %" if md['CI'] != '' else ''),
                 (0.75, 0.25, f"Analysis Version: {self.version}"), (0.75, 0.50, f"Analyzed by: {md['analysisUser']}")]
        for fx, fy, text in lines:
            draw.text((10 + fx * image.size[0], fy * h), text, fill=(255, 255, 255), font=font)
        image.save(path, 'PNG')
        print(f'\033[32mScreenshot saved to {path}\033[37m')

    def _out_of_scope(self, name):
        raise NotImplementedError(
            f"{name} is a file-format exporter outside the N4 -> VDP -> CI hot path (SURVEY.md §2 rows 11-12); "
            "it consumes the arrays this class produces and is unchanged from the reference.")

    def exportNifti(self, filepath=None, fileName=None):
        self._out_of_scope("exportNifti")

    def reconstruct_kspace(self, raw_K):
        """The arithmetic of ``process_RAW`` (Vent_Analysis.py:536-540): per-slice
        ``fftshift(fft2(fftshift(.)))``, then ``transpose(., (1, 0, 2))[:, ::-1, :]`` -> complex128.
        Kernel: ``recon_pass_kernel`` (single precision transforms)."""
        return self._ctx().recon_host([np.asarray(raw_K)])[0].astype(np.complex128)

    def process_RAW(self, filepath=None):
        """Vent_Analysis.py:522-540: read a Siemens twix file (mapVBVD, host side) and reconstruct it."""
        import mapvbvd
        if filepath is None:
            import tkinter.filedialog
            filepath = tkinter.filedialog.askopenfilename()
        self.raw_twix = mapvbvd.mapVBVD(filepath)
        self.metadata['TWIXscanDateTime'] = self.raw_twix.hdr.Config['PrepareTimestamp']
        self.metadata['TWIXprotocolName'] = self.raw_twix.hdr.Meas['tProtocolName']
        self.raw_twix.image.squeeze = True
        self.raw_K = self.raw_twix.image['']
        self.raw_HPvent = self.reconstruct_kspace(self.raw_K)

    def dicom_to_json(self, ds, json_path='c:/pirl/data/DICOMjson.json', include_private=True):
        self._out_of_scope("dicom_to_json")

    def __repr__(self):
        """One line per attribute: arrays by shape, metadata by value (cf. Vent_Analysis.py:562-577)."""
        def show(v):
            if isinstance(v, np.ndarray):
                return f"array{v.shape} {v.dtype}"
            return "(unset)" if isinstance(v, str) and v == '' else (v if np.isscalar(v) else type(v).__name__)
        out = [f"Vent_Analysis (B200) version {self.version}"]
        for name, value in vars(self).items():
            if isinstance(value, dict):
                out += [f"   {k}: {show(v)}" for k, v in value.items()]
            elif not name.startswith('_'):
                out.append(f" {name}: {show(value)}")
        return "\n".join(out)
