"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's MHA ingest for the parity tests of row f1.
Never imported by the product (surffeatures_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU legs.

Follows SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx line by line in plain Python:
  readImageDimensions_mha :444-476, readImageTransforms_mha :478-541, readImages_mha :543-595, read_mha :598-632,
  cropData(cv::Mat&) :745-752, cropRatiosValid :2198-2201.
PARITY UNPINNED: the reference ships no sequence file, test or golden vector for this path; the pins are the
format rules above and files written by surffeatures_b200.mha.write_mha."""
import numpy as np

f32 = np.float32


def _get_dir(filename):
    pos = filename.rfind("/")
    return "" if pos < 0 else filename[:pos] + "/"


def _header_lines(raw):
    """std::getline over the file: lines without their newline, as bytes."""
    return raw.split(b"\n")


def read_image_dimensions(raw):
    for line in _header_lines(raw):
        if line == b"":
            break
        if b"DimSize =" in line:
            parts = line.decode(errors="ignore").split("DimSize =", 1)[1].split()
            if line.startswith(b"DimSize = ") and len(parts) >= 3:
                return int(parts[0]), int(parts[1]), int(parts[2])
            return None
    return None


def _atof(s):
    """C atof: longest numeric prefix after leading blanks, 0.0 if none."""
    s = s.lstrip(" \t")
    best = 0.0
    for end in range(len(s), 0, -1):
        try:
            best = float(s[:end])
            return best
        except ValueError:
            continue
    return best


def read_image_transforms(filename, raw):
    dir_name = _get_dir(filename)
    transforms, validity, names = [], [], []
    for bline in _header_lines(raw):
        if bline == b"":
            break
        line = bline.decode(errors="ignore")
        if "ProbeToTrackerTransform =" in line or "UltrasoundToTrackerTransform =" in line:
            eq = line.index("=")
            name = line[:eq - 1]
            vals = []
            p = eq
            ok = True
            for _ in range(12):
                p = line.find(" ", p + 1)
                if p < 0:
                    ok = False
                    break
                vals.append(f32(_atof(line[p:])))
                p += 1
            if not ok:
                return transforms, validity, names
            transforms.append(vals)
            names.append(dir_name + name + ".png")
        elif "UltrasoundToTrackerTransformStatus" in line or "ProbeToTrackerTransformStatus" in line:
            if "OK" in line:
                validity.append(True)
            elif "INVALID" in line:
                validity.append(False)
        if "ElementDataFile = LOCAL" in line:
            break
    return transforms, validity, names


def read_images(raw, first, last, validity, skip_invalid_bytes=False):
    dims = read_image_dimensions(raw)
    if dims is None:
        return [], first, last
    cols, rows, count = dims
    marker = raw.index(b"ElementDataFile = LOCAL")
    pos = raw.index(b"\n", marker) + 1
    if first < 0:
        first = 0
    if first >= count:
        first = count - 1
    if last < 0 or last >= count:
        last = count - 1
    if first > last:
        first = last
    nbytes = rows * cols
    images = []
    for i in range(count):
        if i < first:
            pos += nbytes
        elif first <= i <= last and validity[i]:
            buf = np.frombuffer(raw[pos:pos + nbytes], np.uint8)
            img = np.zeros(nbytes, np.uint8)
            img[:len(buf)] = buf
            images.append(img.reshape(rows, cols))
            pos += nbytes
        elif i > last:
            break
        elif skip_invalid_bytes:
            pos += nbytes      # NOT in the reference: an INVALID frame leaves the file position where it was
    return images, first, last


def read_mha(filename, first=-1, last=-1, skip_invalid_bytes=False):
    """Returns dict(images, names, transforms, validity, first, last) exactly as read_mha leaves its arguments."""
    raw = open(filename, "rb").read()
    transforms, validity, names = read_image_transforms(filename, raw)
    images, first, last = read_images(raw, first, last, validity, skip_invalid_bytes)
    kept_t, kept_n = [], []
    for i in range(len(validity)):
        if i < first or i > last or not validity[i]:
            continue
        kept_t.append(transforms[i])
        kept_n.append(names[i])
    return dict(images=images, names=kept_n, transforms=np.array(kept_t, np.float32).reshape(-1, 12),
                validity=np.array(validity[first:last + 1], bool), first=first, last=last)


def crop_rect(width, height, ratios):
    """cv::Rect roi(size.width*r0, size.height*r2, size.width*(r1-r0), size.height*(r3-r2)): fp32, truncation."""
    r = [f32(v) for v in ratios]
    x = int(f32(f32(width) * r[0]))
    y = int(f32(f32(height) * r[2]))
    w = int(f32(f32(width) * f32(r[1] - r[0])))
    h = int(f32(f32(height) * f32(r[3] - r[2])))
    return x, y, w, h


def crop(img, ratios):
    x, y, w, h = crop_rect(img.shape[1], img.shape[0], ratios)
    return np.ascontiguousarray(img[y:y + h, x:x + w])
