//! Headless driver: the reference's main() (reference src/main.rs:356-895) without AppKit and Metal.
//! NOT compiled in this environment (no Rust toolchain); the C++ twin mm_headless is built and tested.
mod ffi;
use ffi::*;
use std::{ffi::CStr, io::Write, ptr};

fn check(ctx: *const mm_ctx, what: &str, rc: i32) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(mm_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("{what} failed: {rc} ({msg})");          // the reference panics through expect()/unwrap() too (utils.rs:19,43)
    }
}

fn main() {
    let (maze, w, h, spp, bounces) = (32u32, 1920u32, 1080u32, 16u32, 8u32);
    let gpus: i32 = std::env::args().nth(1).and_then(|a| a.parse().ok()).unwrap_or(1);          // mm_headless_rs [n_gpus]
    if gpus > 1 { return main_multi(maze, w, h, spp, bounces, gpus); }
    let mut scene = ptr::null_mut();
    check(ptr::null(), "mm_scene_build", unsafe { mm_scene_build(maze, 0, 1, &mut scene) });      // main.rs:357-588
    let mut ctx = ptr::null_mut();
    check(ptr::null(), "mm_create", unsafe { mm_create(0, &mut ctx) });                            // main.rs:616-644
    let mut noise = vec![128u8; 512 * 512 * 4];                                                    // texel (0,0) is the only one sampled
    noise.iter_mut().skip(3).step_by(4).for_each(|a| *a = 255);
    check(ctx, "mm_upload_scene", unsafe {                                                         // main.rs:667-695, 723-730
        mm_upload_scene(ctx, mm_scene_planes(scene), mm_scene_n_planes(scene), mm_scene_nodes(scene), mm_scene_n_nodes(scene),
                        mm_scene_indices(scene), mm_scene_materials(scene), mm_scene_emissions(scene), noise.as_ptr(), 512, 512)
    });
    let mut uni = Uniform::default();
    check(ctx, "mm_default_uniform", unsafe { mm_default_uniform(maze, w as f32, h as f32, 4, 0, &mut uni) });   // main.rs:732-755
    let n = unsafe { mm_gen_chunks(w as f32, h as f32, 4, ptr::null_mut(), 0) };
    let mut chunks = vec![Chunk::default(); n as usize];
    unsafe { mm_gen_chunks(w as f32, h as f32, 4, chunks.as_mut_ptr(), n) };                        // main.rs:293-302
    let params = Params { spp, bounce_limit: bounces, mirror_limit: 15, grid_x: w / 4, grid_y: h / 4, ..Default::default() };
    let mut frame = vec![0f32; (w * h * 4) as usize];
    // pin + map the Vec once: the kernel writes finished pixels straight into it (an unregistered Vec works too, through a
    // staged copy in the library)
    check(ctx, "mm_host_register", unsafe { mm_host_register(frame.as_mut_ptr() as *mut _, frame.len() * 4) });
    let mut counters = Counters::default();
    for t in 0..3u32 {
        uni.time = t;                                                                              // main.rs:857
        // the chunk list goes up with the first frame and is kept (null) afterwards: a full-frame list does not change
        let cl = if t == 0 { chunks.as_ptr() } else { ptr::null() };
        check(ctx, "mm_render_async", unsafe {                                                     // main.rs:867-886, commit (:894)
            mm_render_async(ctx, &uni, &params, cl, n, frame.as_mut_ptr(), ptr::null())
        });
        // ... the CPU is free here, as in the reference's loop between commit() and the next drawable ...
        check(ctx, "mm_wait", unsafe { mm_wait(ctx, &mut counters) });
    }
    let mut ms = 0f32;
    unsafe { mm_last_ms(ctx, &mut ms) };
    println!("{} rays, kernel {:.3} ms, {:.1} Mrays/s", counters.rays, ms, counters.rays as f64 / ms as f64 / 1e3);
    let mut f = std::fs::File::create("frame.ppm").unwrap();
    write!(f, "P6\n{} {}\n255\n", w, h).unwrap();
    let bytes: Vec<u8> = frame.chunks(4).flat_map(|p| [p[0], p[1], p[2]]).map(|v| (v.clamp(0.0, 1.0) * 255.0).round() as u8).collect();
    f.write_all(&bytes).unwrap();
    unsafe { mm_host_unregister(frame.as_mut_ptr() as *mut _); mm_destroy(ctx); mm_scene_free(scene); }
}

/// The same frame split over `gpus` devices of this box in ONE process (mm_multi): groups interleaved over the devices,
/// pixels exchanged by the render kernel's NVLink peer stores, frame assembled in the registered Vec by zero-copy stores.
fn main_multi(maze: u32, w: u32, h: u32, spp: u32, bounces: u32, gpus: i32) {
    let mut scene = ptr::null_mut();
    check(ptr::null(), "mm_scene_build", unsafe { mm_scene_build(maze, 0, 1, &mut scene) });
    let devices: Vec<i32> = (0..gpus).collect();
    let mut m = ptr::null_mut();
    let rc = unsafe { mm_multi_create(devices.as_ptr(), gpus, MM_EXCHANGE_PEER, &mut m) };
    if rc != 0 { panic!("mm_multi_create failed: {rc} ({})", unsafe { CStr::from_ptr(mm_multi_last_error(ptr::null())) }.to_string_lossy()); }
    let mut noise = vec![128u8; 512 * 512 * 4];
    noise.iter_mut().skip(3).step_by(4).for_each(|a| *a = 255);
    let ck = |what: &str, rc: i32| if rc != 0 { panic!("{what} failed: {rc} ({})", unsafe { CStr::from_ptr(mm_multi_last_error(m)) }.to_string_lossy()) };
    ck("mm_multi_upload_scene", unsafe {
        mm_multi_upload_scene(m, mm_scene_planes(scene), mm_scene_n_planes(scene), mm_scene_nodes(scene), mm_scene_n_nodes(scene),
                              mm_scene_indices(scene), mm_scene_materials(scene), mm_scene_emissions(scene), noise.as_ptr(), 512, 512)
    });
    let mut uni = Uniform::default();
    unsafe { mm_default_uniform(maze, w as f32, h as f32, 4, 0, &mut uni) };
    let n = unsafe { mm_gen_chunks(w as f32, h as f32, 4, ptr::null_mut(), 0) };
    let mut chunks = vec![Chunk::default(); n as usize];
    unsafe { mm_gen_chunks(w as f32, h as f32, 4, chunks.as_mut_ptr(), n) };
    let params = Params { spp, bounce_limit: bounces, mirror_limit: 15, grid_x: w / 4, grid_y: h / 4, ..Default::default() };
    let mut frame = vec![0f32; (w * h * 4) as usize];
    unsafe { mm_host_register(frame.as_mut_ptr() as *mut _, frame.len() * 4) };
    let mut counters = Counters::default();
    ck("mm_multi_render", unsafe { mm_multi_render(m, &uni, &params, chunks.as_ptr(), n, frame.as_mut_ptr(), &mut counters) });
    let mut ms = 0f32;
    unsafe { mm_multi_last_ms(m, &mut ms) };
    println!("{gpus} GPUs: {} rays, slowest kernel {:.3} ms", counters.rays, ms);
    unsafe { mm_host_unregister(frame.as_mut_ptr() as *mut _); mm_multi_destroy(m); mm_scene_free(scene); }
}
