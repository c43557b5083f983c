//! Writes reference dumps of `TerrainGenerator::generate()` (fastlem 0.1.4) and of `Terrain2D::get_elevation` in the flat
//! little-endian format read by tests/test_reference_dumps.py (format "FLDUMP01", described there and below).
//!
//!     cargo run --release -- <scenario|all> <output directory> [sites]
//!
//! Scenarios (each cites the reference file it repeats):
//!   c1     examples/landscape_evolution.rs:18-34     random sites in [0,100]^2, relaxate_sites(1), erodibility 1.0
//!   slope  tests/landscape_evolution.rs:9-32         [0,200]x[0,100], every setter, max_slope = Some(3.14 * 0.1)
//!   edge   examples/terrain_generation_advanced.rs:36-42 (model part)  ... + add_edge_sites(None, None): equally spaced rim
//!          sites (exact edge-length ties), explicit outlets on a rim band, erodibility varying from site to site
//! Only the crate's public API is used: the model (sites, areas, default outlets, the graph walked through
//! `neighbors_of`, i.e. in the adjacency order every tie-break of the solver depends on), the parameters, and the
//! elevations `generate()` returns for max_iteration = 1, 2, 5, 20 and "until stable".  The stream tree itself is private
//! to the crate, but the elevations after ONE iteration already depend on every stage (receivers, lake removal, drainage
//! areas, response times, the slope clamp), so they pin the whole loop body.
use fastlem::core::parameters::TopographicalParameters;
use fastlem::core::traits::Model;
use fastlem::lem::generator::TerrainGenerator;
use fastlem::models::surface::model::TerrainModel2D;
use fastlem::models::surface::{builder::TerrainModel2DBulider, sites::Site2D};
use std::fs::File;
use std::io::{BufWriter, Write};

const SNAPSHOTS: [Option<u32>; 5] = [Some(1), Some(2), Some(5), Some(20), None];
const QUERY_SIDE: usize = 64;

struct Params {
    base: Vec<f64>,
    erodibility: Vec<f64>,
    uplift: Vec<f64>,
    max_slope: Vec<Option<f64>>,
    is_outlet: Vec<bool>,
}

impl Params {
    fn to_vec(&self) -> Vec<TopographicalParameters> {
        (0..self.base.len())
            .map(|i| {
                TopographicalParameters::default()
                    .set_base_elevation(self.base[i])
                    .set_erodibility(self.erodibility[i])
                    .set_uplift_rate(self.uplift[i])
                    .set_is_outlet(self.is_outlet[i])
                    .set_max_slope(self.max_slope[i])
            })
            .collect()
    }
}

fn put_u32(w: &mut impl Write, v: &[u32]) {
    for x in v {
        w.write_all(&x.to_le_bytes()).unwrap();
    }
}
fn put_u64(w: &mut impl Write, v: &[u64]) {
    for x in v {
        w.write_all(&x.to_le_bytes()).unwrap();
    }
}
fn put_f64(w: &mut impl Write, v: &[f64]) {
    for x in v {
        w.write_all(&x.to_le_bytes()).unwrap();
    }
}

fn dump(path: &str, model: &TerrainModel2D, bound_min: Site2D, bound_max: Site2D, params: &Params) {
    let n = model.num();
    let graph = model.graph();
    // CSR in neighbors_of order
    let mut row_ptr: Vec<u32> = vec![0; n + 1];
    let mut col: Vec<u32> = Vec::new();
    let mut dist: Vec<f64> = Vec::new();
    for i in 0..n {
        for ja in graph.neighbors_of(i).iter() {
            col.push(ja.0 as u32);
            dist.push(ja.1);
        }
        row_ptr[i + 1] = col.len() as u32;
    }
    let outlets: Vec<u32> = model.default_outlets().iter().map(|&i| i as u32).collect();
    // query points of the get_elevation leg: pixel centres of a QUERY_SIDE^2 raster over the bounding box
    // (examples/terrain_generation_advanced.rs:296-299)
    let mut queries: Vec<f64> = Vec::new();
    for row in 0..QUERY_SIDE {
        for c in 0..QUERY_SIDE {
            queries.push(bound_min.x + (bound_max.x - bound_min.x) * ((c as f64 + 0.5) / QUERY_SIDE as f64));
            queries.push(bound_min.y + (bound_max.y - bound_min.y) * ((row as f64 + 0.5) / QUERY_SIDE as f64));
        }
    }

    let mut w = BufWriter::new(File::create(path).unwrap());
    w.write_all(b"FLDUMP01").unwrap();
    put_u64(&mut w, &[n as u64, col.len() as u64, outlets.len() as u64, SNAPSHOTS.len() as u64, (queries.len() / 2) as u64]);
    put_f64(&mut w, &[bound_min.x, bound_min.y, bound_max.x, bound_max.y]);
    let sites: Vec<f64> = model.sites().iter().flat_map(|s| [s.x, s.y]).collect();
    put_f64(&mut w, &sites);
    put_f64(&mut w, model.areas());
    put_u32(&mut w, &row_ptr);
    put_u32(&mut w, &col);
    put_f64(&mut w, &dist);
    put_u32(&mut w, &outlets);
    put_f64(&mut w, &params.base);
    put_f64(&mut w, &params.erodibility);
    put_f64(&mut w, &params.uplift);
    let ms: Vec<f64> = params.max_slope.iter().map(|m| m.unwrap_or(f64::NAN)).collect();
    put_f64(&mut w, &ms);
    let io: Vec<u8> = params.is_outlet.iter().map(|&b| b as u8).collect();
    w.write_all(&io).unwrap();

    let mut last = None;
    for snap in SNAPSHOTS.iter() {
        let mut generator = TerrainGenerator::default().set_model(model.clone()).set_parameters(params.to_vec());
        if let Some(k) = snap {
            generator = generator.set_max_iteration(*k);
        }
        let terrain = generator.generate().unwrap();
        put_u32(&mut w, &[snap.unwrap_or(u32::MAX)]);
        put_f64(&mut w, terrain.elevations());
        last = Some(terrain);
    }
    // Terrain2D::get_elevation of the converged terrain (None -> NaN)
    let terrain = last.unwrap();
    let values: Vec<f64> = queries
        .chunks(2)
        .map(|q| terrain.get_elevation(&Site2D { x: q[0], y: q[1] }).unwrap_or(f64::NAN))
        .collect();
    put_f64(&mut w, &queries);
    put_f64(&mut w, &values);
    w.flush().unwrap();
    println!("wrote {} ({} sites, {} directed edges)", path, n, col.len());
}

fn uniform(n: usize) -> Params {
    Params { base: vec![0.0; n], erodibility: vec![1.0; n], uplift: vec![1.0; n], max_slope: vec![None; n], is_outlet: vec![false; n] }
}

fn scenario(name: &str, dir: &str, sites: Option<usize>) {
    match name {
        "c1" => {
            let num = sites.unwrap_or(30000);
            let (lo, hi) = (Site2D { x: 0.0, y: 0.0 }, Site2D { x: 100.0, y: 100.0 });
            let model = TerrainModel2DBulider::from_random_sites(num, lo, hi).relaxate_sites(1).unwrap().build().unwrap();
            let n = model.num();
            dump(&format!("{}/ref_c1_{}.bin", dir, num), &model, lo, hi, &uniform(n));
        }
        "slope" => {
            let num = sites.unwrap_or(10000);
            let (lo, hi) = (Site2D { x: 0.0, y: 0.0 }, Site2D { x: 200.0, y: 100.0 });
            let model = TerrainModel2DBulider::from_random_sites(num, lo, hi).relaxate_sites(1).unwrap().build().unwrap();
            let n = model.num();
            let mut p = uniform(n);
            p.max_slope = vec![Some(3.14 * 0.1); n];
            dump(&format!("{}/ref_slope_{}.bin", dir, num), &model, lo, hi, &p);
        }
        "edge" => {
            let num = sites.unwrap_or(20000);
            let (lo, hi) = (Site2D { x: 0.0, y: 0.0 }, Site2D { x: 100.0, y: 100.0 });
            let model = TerrainModel2DBulider::from_random_sites(num, lo, hi)
                .relaxate_sites(1)
                .unwrap()
                .add_edge_sites(None, None)
                .unwrap()
                .build()
                .unwrap();
            let n = model.num();
            let mut p = uniform(n);
            for (i, s) in model.sites().iter().enumerate() {
                // a rim band of explicit outlets (the tied edge sites included) and a site-to-site varying erodibility
                p.is_outlet[i] = s.x < 4.0 || s.y < 4.0 || s.x > 96.0 || s.y > 96.0;
                p.erodibility[i] = 0.5 + ((i as u64).wrapping_mul(2654435761) % 1000) as f64 / 1000.0;
            }
            dump(&format!("{}/ref_edge_{}.bin", dir, num), &model, lo, hi, &p);
        }
        other => panic!("unknown scenario {}", other),
    }
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    if args.len() < 3 {
        eprintln!("usage: fastlem_dump <c1|slope|edge|all> <output directory> [sites]");
        std::process::exit(2);
    }
    let sites = args.get(3).map(|s| s.parse::<usize>().unwrap());
    if args[1] == "all" {
        for s in ["c1", "slope", "edge"] {
            scenario(s, &args[2], sites);
        }
    } else {
        scenario(&args[1], &args[2], sites);
    }
}
