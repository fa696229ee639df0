/* Build stub: the reference's draw.hh includes <cairo/cairo.h> only to name two
 * opaque types in declarations. libcairo is not available offline, and the
 * renderer (draw.cpp) is not compiled into oracle/_ref; see oracle/README.md. */
typedef struct _cairo_surface cairo_surface_t;
typedef struct _cairo cairo_t;
