//! `GpuFilterTaskBuilder` / `GpuMaterializeFilesTaskBuilder`: the reference's producer tasks with the compute call
//! swapped for the C-ABI (feature "chapterhouse-tasks": this file compiles inside the ChapterhouseDB crate tree,
//! next to operators/filter_tasks and operators/materialize_tasks, whose types it uses unchanged).
//!
//! Control flow is filter_tasks/filter_task.rs:65-142 and materialize_tasks/materialize_files_task.rs:68-170 line
//! for line -- pull from the inbound exchange, compute, push / write, ack -- so the exchange protocol, heartbeats,
//! cancellation and error propagation (`?` -> `tx.send(Some(err))` -> OperatorInstanceStatusChange::Error) are the
//! reference's own.  What changes:
//!   * filter_task.rs:99-103   `record_utils::filter_record(rec, aliases, &expr)?`
//!       -> program compiled once per (schema, aliases) with `chdb_gpu::compile_filter`, then `chdb_gpu::run_async`
//!          (non-blocking: the tokio worker is not parked on the GPU);
//!   * materialize_files_task.rs:110-114 `record_utils::project_record(&fields, rec, aliases)?` likewise;
//!   * one `GpuContext` per operator instance: device = operator_instance index mod visible GPUs (worker config
//!     `gpus`), so N instances of one operator drain one exchange queue on N GPUs.
//! Registration (operator_task_registry.rs:150-162):
//!     OperatorTaskRegistry::new()
//!         .add_table_func_task_builder(Box::new(ReadFilesTaskBuilder::new()), Box::new(ReadFilesSyntaxValidator::new()))?
//!         .add_filter_task_builder(Box::new(chdb_gpu::tasks::GpuFilterTaskBuilder::new(gpus)))?
//!         .add_materialize_files_builder(Box::new(chdb_gpu::tasks::GpuMaterializeFilesTaskBuilder::new(gpus)), vec![DataFormat::Parquet])?
use std::{path::PathBuf, sync::Arc};

use anyhow::{Error, Result};
use tokio::sync::Mutex;
use tokio_util::sync::CancellationToken;
use tracing::{debug, error};
use uuid::Uuid;

use crate::handlers::exchange_handlers;
use crate::handlers::message_router_handler::{MessageConsumer, MessageRouterState};
use crate::handlers::{
    message_handler::{MessageRegistry, Pipe},
    operator_handler::{
        operator_handler_state::OperatorInstanceConfig,
        operators::{
            filter_tasks::{config::FilterConfig, filter_task::FilterConsumer},
            materialize_tasks::{config::MaterializeFilesConfig, materialize_files_task::MaterializeFilesConsumer},
            operator_task_trackers::RestrictedOperatorTaskTracker,
            traits::TaskBuilder,
            ConnectionRegistry,
        },
    },
};
use crate::{compile_filter, compile_project, run_async, GpuContext, GpuProgram};

/// Which GPU an operator instance runs on: instances of one operator are spread round-robin over the worker's GPUs.
fn device_for(op_in_config: &OperatorInstanceConfig, gpus: i32) -> i32 {
    (op_in_config.id % gpus.max(1) as u128) as i32
}

/// The program of the instance's expression for the schema of the records it sees (compiled on the first record;
/// recompiled only if a later record has another schema or alias list).
struct ProgramCache {
    key: Option<(arrow::datatypes::SchemaRef, Vec<Vec<String>>)>,
    prog: Option<GpuProgram>,
}

impl ProgramCache {
    fn new() -> Self {
        ProgramCache { key: None, prog: None }
    }
    fn get<F>(&mut self, rec: &arrow::array::RecordBatch, aliases: &Vec<Vec<String>>, compile: F) -> Result<&GpuProgram>
    where
        F: FnOnce() -> Result<GpuProgram>,
    {
        let same = matches!(&self.key, Some((s, a)) if s == &rec.schema() && a == aliases);
        if !same {
            self.prog = Some(compile()?);
            self.key = Some((rec.schema(), aliases.clone()));
        }
        Ok(self.prog.as_ref().expect("compiled above"))
    }
}

///////////////////////////////////////////////////////
// GPU filter producer

struct GpuFilterTask {
    operator_instance_config: OperatorInstanceConfig,
    filter_config: FilterConfig,
    operator_pipe: Pipe,
    msg_reg: Arc<MessageRegistry>,
    msg_router_state: Arc<Mutex<MessageRouterState>>,
    ctx: GpuContext,
}

impl GpuFilterTask {
    async fn async_main(&mut self, ct: CancellationToken) -> Result<()> {
        let mut rec_handler = exchange_handlers::record_handler::RecordHandler::initiate(
            ct.child_token(),
            &self.operator_instance_config,
            &mut self.operator_pipe,
            self.msg_reg.clone(),
            self.msg_router_state.clone(),
        )
        .await?;
        let mut programs = ProgramCache::new();

        loop {
            let exchange_rec = rec_handler
                .next_record(ct.child_token(), &mut self.operator_pipe, None)
                .await?;
            match exchange_rec {
                Some(exchange_rec) => {
                    debug!(
                        record_id = exchange_rec.record_id,
                        record_num_rows = exchange_rec.record.num_rows(),
                        device = self.ctx.device(),
                        "received record"
                    );
                    // filter the record (filter_task.rs:99-103) on this instance's GPU
                    let expr = &self.filter_config.expr;
                    let prog = programs.get(&exchange_rec.record, &exchange_rec.table_aliases, || {
                        compile_filter(&exchange_rec.record, &exchange_rec.table_aliases, expr)
                    })?;
                    let filtered_rec = run_async(&self.ctx, prog, exchange_rec.record.clone()).await?;

                    // send the record to the outbound exchange
                    rec_handler
                        .send_record_to_outbound_exchange(
                            &mut self.operator_pipe,
                            exchange_rec.record_id.clone(),
                            filtered_rec,
                            exchange_rec.table_aliases.clone(),
                        )
                        .await?;
                    // confirm processing of the record with the inbound exchange
                    rec_handler.complete_record(&mut self.operator_pipe, exchange_rec).await?;
                }
                None => {
                    debug!("read all records from the exchange");
                    break;
                }
            }
        }
        if let Err(err) = rec_handler.close().await {
            error!("{}", err);
        }
        Ok(())
    }
}

#[derive(Debug, Clone)]
pub struct GpuFilterTaskBuilder {
    gpus: i32,
}

impl GpuFilterTaskBuilder {
    pub fn new(gpus: i32) -> GpuFilterTaskBuilder {
        GpuFilterTaskBuilder { gpus }
    }
}

impl TaskBuilder for GpuFilterTaskBuilder {
    fn build(
        &self,
        op_in_config: OperatorInstanceConfig,
        operator_pipe: Pipe,
        msg_reg: Arc<MessageRegistry>,
        _conn_reg: Arc<ConnectionRegistry>,
        msg_router_state: Arc<Mutex<MessageRouterState>>,
        tt: &mut RestrictedOperatorTaskTracker,
        ct: CancellationToken,
    ) -> Result<(tokio::sync::oneshot::Receiver<Option<Error>>, Box<dyn MessageConsumer>)> {
        let filter_config = FilterConfig::try_from(&op_in_config)?;
        let ctx = GpuContext::new(device_for(&op_in_config, self.gpus))?;
        let consumer: Box<dyn MessageConsumer> = Box::new(FilterConsumer::new(msg_reg.clone()));
        let mut op = GpuFilterTask {
            operator_instance_config: op_in_config,
            filter_config,
            operator_pipe,
            msg_reg,
            msg_router_state,
            ctx,
        };
        let (tx, rx) = tokio::sync::oneshot::channel();
        tt.spawn(async move {
            let res = op.async_main(ct).await;
            if let Err(ref err) = res {
                error!("{:?}", err);
            }
            if let Err(err_send) = tx.send(res.err()) {
                error!("{:?}", err_send);
            }
        })?;
        Ok((rx, consumer))
    }
}

///////////////////////////////////////////////////////
// GPU materialize producer

struct GpuMaterializeFilesTask {
    operator_instance_config: OperatorInstanceConfig,
    materialize_file_config: MaterializeFilesConfig,
    operator_pipe: Pipe,
    msg_reg: Arc<MessageRegistry>,
    conn_reg: Arc<ConnectionRegistry>,
    msg_router_state: Arc<Mutex<MessageRouterState>>,
    ctx: GpuContext,
}

impl GpuMaterializeFilesTask {
    /// The compaction the reference lists as a TODO (DEV_NOTES.md:117-122), on the device: projected records stay in HBM
    /// until `max_rows_per_row_group * max_row_groups` rows are pending (or the exchange is drained), then ONE call encodes
    /// them into a `part_<n>.parquet` image in pinned memory (`chdb_parquet_encode`: a page per record and column, records
    /// coalesced into row groups) and only the image crosses PCIe.  Records are completed after their file is written.
    async fn write_compacted(
        &mut self,
        storage_conn: &opendal::Operator,
        query_uuid_id: &Uuid,
        pending: &mut Vec<(crate::DeviceBatch, exchange_handlers::record_handler::ExchangeRecord)>,
        file_no: &mut u64,
        rec_handler: &mut exchange_handlers::record_handler::RecordHandler,
        max_rows_per_row_group: i64,
        max_row_groups: i32,
    ) -> Result<()> {
        while !pending.is_empty() {
            let image = {
                let refs: Vec<&crate::DeviceBatch> = pending.iter().map(|(b, _)| b).collect();
                crate::encode_parquet(&self.ctx, &refs, max_rows_per_row_group, max_row_groups)?
            };
            let path = format!("/query_results/{}/part_{}.parquet", query_uuid_id, *file_no);
            let mut writer = storage_conn.writer_with(&path).chunk(16 * 1024 * 1024).concurrent(4).await?;
            writer.write(image.as_bytes().to_vec()).await?;
            writer.close().await?;
            for (_, rec) in pending.drain(..image.consumed) {
                rec_handler.complete_record(&mut self.operator_pipe, rec).await?;
            }
            *file_no += 1;
        }
        Ok(())
    }

    async fn async_main(&mut self, ct: CancellationToken) -> Result<()> {
        let storage_conn = self.conn_reg.get_operator("default")?;
        let query_uuid_id = Uuid::from_u128(self.operator_instance_config.query_id.clone());
        let mut rec_handler = exchange_handlers::record_handler::RecordHandler::initiate(
            ct.child_token(),
            &self.operator_instance_config,
            &mut self.operator_pipe,
            self.msg_reg.clone(),
            self.msg_router_state.clone(),
        )
        .await?;
        let mut programs = ProgramCache::new();

        loop {
            let exchange_rec = rec_handler
                .next_record(ct.child_token(), &mut self.operator_pipe, None)
                .await?;
            match exchange_rec {
                Some(exchange_rec) => {
                    // project the record (materialize_files_task.rs:110-114) on this instance's GPU
                    let fields = &self.materialize_file_config.fields;
                    let prog = programs.get(&exchange_rec.record, &exchange_rec.table_aliases, || {
                        compile_project(fields, &exchange_rec.record, &exchange_rec.table_aliases)
                    })?;
                    let proj_rec = run_async(&self.ctx, prog, exchange_rec.record.clone()).await?;

                    // materialize the projected record: same path and writer as the reference (:116-141)
                    let mut rec_path_buf = PathBuf::from("/query_results");
                    rec_path_buf.push(format!("{}", query_uuid_id));
                    rec_path_buf.push(format!("rec_{}.parquet", exchange_rec.record_id));
                    let rec_path = rec_path_buf
                        .to_str()
                        .ok_or_else(|| anyhow::anyhow!("record path formatting returned None result"))?
                        .to_string();
                    let writer = storage_conn.writer_with(&rec_path).chunk(16 * 1024 * 1024).concurrent(4).await?;
                    let parquet_writer = parquet_opendal::AsyncWriter::new(writer);
                    let mut arrow_parquet_writer =
                        parquet::arrow::AsyncArrowWriter::try_new(parquet_writer, proj_rec.schema(), None)?;
                    arrow_parquet_writer.write(&proj_rec).await?;
                    arrow_parquet_writer.close().await?;

                    rec_handler.complete_record(&mut self.operator_pipe, exchange_rec).await?;
                }
                None => {
                    debug!("complete materialization; read all records from the exchange");
                    break;
                }
            }
        }
        if let Err(err) = rec_handler.close().await {
            error!("{}", err);
        }
        Ok(())
    }
}

#[derive(Debug, Clone)]
pub struct GpuMaterializeFilesTaskBuilder {
    gpus: i32,
}

impl GpuMaterializeFilesTaskBuilder {
    pub fn new(gpus: i32) -> GpuMaterializeFilesTaskBuilder {
        GpuMaterializeFilesTaskBuilder { gpus }
    }
}

impl TaskBuilder for GpuMaterializeFilesTaskBuilder {
    fn build(
        &self,
        op_in_config: OperatorInstanceConfig,
        operator_pipe: Pipe,
        msg_reg: Arc<MessageRegistry>,
        conn_reg: Arc<ConnectionRegistry>,
        msg_router_state: Arc<Mutex<MessageRouterState>>,
        tt: &mut RestrictedOperatorTaskTracker,
        ct: CancellationToken,
    ) -> Result<(tokio::sync::oneshot::Receiver<Option<Error>>, Box<dyn MessageConsumer>)> {
        let materialize_file_config = MaterializeFilesConfig::try_from(&op_in_config)?;
        let ctx = GpuContext::new(device_for(&op_in_config, self.gpus))?;
        let consumer: Box<dyn MessageConsumer> = Box::new(MaterializeFilesConsumer::new(msg_reg.clone()));
        let mut op = GpuMaterializeFilesTask {
            operator_instance_config: op_in_config,
            materialize_file_config,
            operator_pipe,
            msg_reg,
            conn_reg,
            msg_router_state,
            ctx,
        };
        let (tx, rx) = tokio::sync::oneshot::channel();
        tt.spawn(async move {
            let res = op.async_main(ct).await;
            if let Err(ref err) = res {
                error!("{:?}", err);
            }
            if let Err(err_send) = tx.send(res.err()) {
                error!("{:?}", err_send);
            }
        })?;
        Ok((rx, consumer))
    }
}
