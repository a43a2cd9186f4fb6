// Drop-in replacement for LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs that forwards the per-frame call to
// libmobicuda.so (include/mobicuda.h) through P/Invoke.  SOURCE-ONLY deliverable: this repository's build image has no
// .NET toolchain, so the file is not compiled here; the identical call sequence is exercised by
// mobiclipdecoder_b200/decoder.py (ctypes) in tests/test_gpu_parity.py.
//
// What stays the same for callers (MobiConverter/Program.cs:64-71, 220-252; MobiclipDecoder/Form1.cs:240-302):
//   var d = new MobiclipDecoder(Width, Height, MobiclipDecoder.MobiclipVersion.Moflex3DS);
//   d.Data = frameBytes; d.Offset = 0; Bitmap b = d.DecodeFrame();   // null on any decode error
//   d.Offset (bytes consumed, audio follows at Offset-2), d.Y[0], d.UV[0], d.Stride, d.Quantizer, d.YuvFormat
// What changes: Y[1..5]/UV[1..5] (older reference pictures) stay on the GPU and read as null here; the VLC tables and
// the Internal[] scratch array are no longer public fields (they were implementation details of the C# loops).
using System;
using System.Drawing;
using System.Drawing.Imaging;
using System.Runtime.InteropServices;

namespace LibMobiclip.Codec.Mobiclip
{
    public unsafe class MobiclipDecoder : IDisposable
    {
        const string Lib = "mobicuda";   // libmobicuda.so / mobicuda.dll

        [DllImport(Lib)] static extern int mobi_create(uint width, uint height, int version, int device, out IntPtr handle);
        [DllImport(Lib)] static extern void mobi_destroy(IntPtr handle);
        [DllImport(Lib)] static extern int mobi_decode_frame(IntPtr handle, byte* data, int len, ref int offset);
        [DllImport(Lib)] static extern int mobi_read_planes_strided(IntPtr handle, byte* y, byte* uv);
        [DllImport(Lib)] static extern int mobi_read_bgra(IntPtr handle, byte* dst, int dstStride);
        [DllImport(Lib)] static extern int mobi_get_state(IntPtr handle, out uint quantizer, out uint yuvFormat, out int stride);
        [DllImport(Lib)] static extern IntPtr mobi_last_error(IntPtr handle);

        public byte[] Data;
        public int Offset = 0;
        public uint Width;
        public uint Height;
        public byte[][] Y = new byte[6][];
        public byte[][] UV = new byte[6][];
        public uint Quantizer = 0;
        public uint YuvFormat;
        public int Stride = 512;

        public enum MobiclipVersion { VxDS, ModsDS, Moflex3DS }
        public MobiclipVersion Version;

        /// <summary>CUDA ordinal the decoder's pictures live on; one decoder = one stream = one device.</summary>
        public static int Device = 0;
        /// <summary>Set to false when the caller only needs the Bitmap (skips the plane read-back).</summary>
        public bool ReadPlanes = true;

        IntPtr handle;

        public MobiclipDecoder(uint Width, uint Height, MobiclipVersion Version)
        {
            this.Width = Width;
            this.Height = Height;
            if (Width <= 256) Stride = 256;
            else if (Width <= 512) Stride = 512;
            else Stride = 1024;
            this.Version = Version;
            // VxDS: DecodeVXS1 throws NotImplementedException in the reference; here creation reports "unsupported"
            // and DecodeFrame returns null.
            if (mobi_create(Width, Height, (int)Version, Device, out handle) != 0) handle = IntPtr.Zero;
        }

        public Bitmap DecodeFrame()
        {
            if (handle == IntPtr.Zero || Data == null) return null;
            int rc;
            fixed (byte* p = Data) rc = mobi_decode_frame(handle, p, Data.Length, ref Offset);
            if (rc != 0) return null;   // the reference swallows every exception and returns null (MobiclipDecoder.cs:325-328)
            int s;
            mobi_get_state(handle, out Quantizer, out YuvFormat, out s);
            if (ReadPlanes)
            {
                Y[0] = new byte[Stride * Height];
                UV[0] = new byte[Stride * Height / 2];
                fixed (byte* y = Y[0]) fixed (byte* uv = UV[0]) mobi_read_planes_strided(handle, y, uv);
            }
            Bitmap b = new Bitmap((int)Width, (int)Height);
            BitmapData d = b.LockBits(new Rectangle(0, 0, b.Width, b.Height), ImageLockMode.WriteOnly, PixelFormat.Format32bppArgb);
            mobi_read_bgra(handle, (byte*)d.Scan0, d.Stride);
            b.UnlockBits(d);
            return b;
        }

        public string LastError { get { return handle == IntPtr.Zero ? "no decoder" : Marshal.PtrToStringAnsi(mobi_last_error(handle)); } }

        public void Dispose()
        {
            if (handle != IntPtr.Zero) { mobi_destroy(handle); handle = IntPtr.Zero; }
            GC.SuppressFinalize(this);
        }
        ~MobiclipDecoder() { if (handle != IntPtr.Zero) mobi_destroy(handle); }
    }
}
