// Oracle tooling only: inert stand-in for <cairo/cairo.h> (Cairo is absent from this image). Every drawing
// call is a no-op and nothing is written to disk; it exists so that the reference's unit tests, which plot
// their results as a side effect, run to completion. Plots are not part of the reachability path.
#ifndef ORACLE_SHIM_NULLCAIRO_H
#define ORACLE_SHIM_NULLCAIRO_H
struct cairo_t { double line_width; };
struct cairo_surface_t { int w; int h; };
struct cairo_text_extents_t { double x_bearing, y_bearing, width, height, x_advance, y_advance; };
enum cairo_format_t { CAIRO_FORMAT_ARGB32=0, CAIRO_FORMAT_RGB24=1 };
enum cairo_font_slant_t { CAIRO_FONT_SLANT_NORMAL=0 };
enum cairo_font_weight_t { CAIRO_FONT_WEIGHT_NORMAL=0 };
enum cairo_status_t { CAIRO_STATUS_SUCCESS=0 };
inline cairo_surface_t* cairo_image_surface_create(cairo_format_t, int w, int h) { cairo_surface_t* s=new cairo_surface_t; s->w=w; s->h=h; return s; }
inline cairo_surface_t* cairo_image_surface_create_for_data(unsigned char*, cairo_format_t, int w, int h, int) { return cairo_image_surface_create(CAIRO_FORMAT_ARGB32,w,h); }
inline int cairo_format_stride_for_width(cairo_format_t, int w) { return 4*w; }
inline void cairo_surface_destroy(cairo_surface_t* s) { delete s; }
inline cairo_status_t cairo_surface_write_to_png(cairo_surface_t*, const char*) { return CAIRO_STATUS_SUCCESS; }
inline cairo_t* cairo_create(cairo_surface_t*) { cairo_t* c=new cairo_t; c->line_width=2.0; return c; }
inline void cairo_destroy(cairo_t* c) { delete c; }
inline void cairo_move_to(cairo_t*, double, double) { }
inline void cairo_line_to(cairo_t*, double, double) { }
inline void cairo_arc(cairo_t*, double, double, double, double, double) { }
inline void cairo_text_extents(cairo_t*, const char*, cairo_text_extents_t* e) { e->x_bearing=e->y_bearing=0; e->width=e->height=10; e->x_advance=e->y_advance=10; }
inline void cairo_show_text(cairo_t*, const char*) { }
inline void cairo_set_source_rgb(cairo_t*, double, double, double) { }
inline void cairo_set_line_width(cairo_t* c, double w) { c->line_width=w; }
inline double cairo_get_line_width(cairo_t* c) { return c->line_width; }
inline void cairo_translate(cairo_t*, double, double) { }
inline void cairo_scale(cairo_t*, double, double) { }
inline void cairo_stroke(cairo_t*) { }
inline void cairo_clip(cairo_t*) { }
inline void cairo_set_font_size(cairo_t*, double) { }
inline void cairo_select_font_face(cairo_t*, const char*, cairo_font_slant_t, cairo_font_weight_t) { }
inline void cairo_save(cairo_t*) { }
inline void cairo_restore(cairo_t*) { }
inline void cairo_paint(cairo_t*) { }
inline void cairo_new_path(cairo_t*) { }
inline void cairo_fill_preserve(cairo_t*) { }
#endif
