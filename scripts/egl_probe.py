"""Probe: can a headless EGL + OpenGL context be created on this box with libEGL_nvidia.so.0?"""
import ctypes as C, sys
try:
    egl = C.CDLL("libEGL_nvidia.so.0")
except OSError as e:
    print("no libEGL_nvidia", e); sys.exit(0)
names = [n for n in ("eglGetProcAddress", "eglInitialize", "eglGetDisplay", "eglChooseConfig", "eglCreateContext",
                     "eglMakeCurrent", "eglBindAPI", "eglQueryString", "eglGetError", "__egl_Main") if hasattr(egl, n)]
print("exports:", names)
if not hasattr(egl, "eglGetProcAddress"):
    sys.exit(0)
gpa = egl.eglGetProcAddress; gpa.restype = C.c_void_p; gpa.argtypes = [C.c_char_p]
def fn(name, restype, *argtypes):
    p = gpa(name.encode())
    if not p:
        p = C.cast(getattr(egl, name), C.c_void_p).value if hasattr(egl, name) else None
    if not p: raise RuntimeError("missing " + name)
    return C.CFUNCTYPE(restype, *argtypes)(p)
EGLDisplay = C.c_void_p
queryDevices = fn("eglQueryDevicesEXT", C.c_uint, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int))
getPlatformDisplay = fn("eglGetPlatformDisplayEXT", EGLDisplay, C.c_uint, C.c_void_p, C.POINTER(C.c_int))
devs = (C.c_void_p * 16)(); n = C.c_int()
print("queryDevices", queryDevices(16, devs, C.byref(n)), n.value)
dpy = getPlatformDisplay(0x313F, devs[0], None)   # EGL_PLATFORM_DEVICE_EXT
print("display", dpy)
egl.eglInitialize.argtypes = [EGLDisplay, C.POINTER(C.c_int), C.POINTER(C.c_int)]
maj, mnr = C.c_int(), C.c_int()
print("init", egl.eglInitialize(dpy, C.byref(maj), C.byref(mnr)), maj.value, mnr.value)
egl.eglBindAPI.argtypes = [C.c_uint]
print("bindAPI(OpenGL)", egl.eglBindAPI(0x30A2))
cfg_attr = (C.c_int * 5)(0x3033, 0x0001, 0x3040, 0x0008, 0x3038)  # SURFACE_TYPE PBUFFER, RENDERABLE_TYPE OPENGL_BIT, NONE
cfg = C.c_void_p(); ncfg = C.c_int()
egl.eglChooseConfig.argtypes = [EGLDisplay, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int)]
print("chooseConfig", egl.eglChooseConfig(dpy, cfg_attr, C.byref(cfg), 1, C.byref(ncfg)), ncfg.value)
ctx_attr = (C.c_int * 5)(0x3098, 4, 0x30FB, 5, 0x3038)   # MAJOR 4, MINOR 5
egl.eglCreateContext.restype = C.c_void_p
egl.eglCreateContext.argtypes = [EGLDisplay, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
ctx = egl.eglCreateContext(dpy, cfg, None, ctx_attr)
print("context", ctx, hex(egl.eglGetError()))
egl.eglMakeCurrent.argtypes = [EGLDisplay, C.c_void_p, C.c_void_p, C.c_void_p]
print("makeCurrent (surfaceless)", egl.eglMakeCurrent(dpy, None, None, ctx), hex(egl.eglGetError()))
glGetString = fn("glGetString", C.c_char_p, C.c_uint)
print("GL_VERSION", glGetString(0x1F02), "GL_RENDERER", glGetString(0x1F01))
